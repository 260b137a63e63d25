"""GPU parity of the individual kernels (through the C ABI) against the fp32 oracle / plain fp32 torch math.
Tolerances: bf16 operands + fp32 accumulation vs fp32 reference -> rel-L2 <= 2e-2, cosine >= 0.999;
fp32-in/fp32-out kernels (LayerNorm fp32 output, loss, normalise) -> 1e-4; index ops bit exact."""
import pytest
import torch

from oracle import dual_encoder as O
from tests.util import assert_close_bf16, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from lavila_b200 import ops, _lib
    return ops, _lib


@pytest.mark.parametrize("M,N,K", [(304, 512, 256), (128, 256, 64), (1000, 768, 3072)])
def test_gemm_layouts(M, N, K):
    ops, L = _ops()
    torch.manual_seed(0)
    A = torch.randn(M, K, device=DEV).bfloat16()
    B = torch.randn(N, K, device=DEV).bfloat16()
    ref = A.float() @ B.float().t()
    out = torch.empty(M, N, device=DEV)
    ops.gemm(A, B, M, N, K, out)
    assert rel_l2(out, ref) < 1e-5
    out = torch.empty(M, N, device=DEV)
    ops.gemm(A, B.t().contiguous(), M, N, K, out, b_mn=1)
    assert rel_l2(out, ref) < 1e-5
    out = torch.zeros(M, N, device=DEV)
    ops.gemm(A.t().contiguous(), B.t().contiguous(), M, N, K, out, a_mn=1, b_mn=1, flags=L.EPI_ATOMIC, k_splits=3)
    assert rel_l2(out, ref) < 1e-5
    outb = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, M, N, K, outb)
    assert rel_l2(outb, ref) < 4e-3


def test_gemm_epilogues():
    ops, L = _ops()
    torch.manual_seed(1)
    M, N, K = 777, 512, 384
    A = (torch.randn(M, K, device=DEV) * 0.2).bfloat16()
    B = (torch.randn(N, K, device=DEV) * 0.2).bfloat16()
    bias, resid = torch.randn(N, device=DEV), torch.randn(M, N, device=DEV)
    aux = torch.randn(M, N, device=DEV).bfloat16()
    alpha = torch.tensor(0.7, device=DEV)
    base = A.float() @ B.float().t() + bias
    act = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    pre = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, M, N, K, act, flags=L.EPI_BIAS | L.EPI_QUICKGELU, bias=bias, out2=pre)
    hb = base.bfloat16().float()
    assert rel_l2(pre, base) < 4e-3
    assert rel_l2(act, O.quick_gelu(hb)) < 6e-3
    dh = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, B, M, N, K, dh, flags=L.EPI_DQUICKGELU, aux=aux)
    h = aux.float().requires_grad_(True)
    (gref,) = torch.autograd.grad(O.quick_gelu(h).sum(), h)
    assert rel_l2(dh, (A.float() @ B.float().t()) * gref) < 6e-3
    y = torch.empty(M, N, device=DEV)
    ops.gemm(A, B, M, N, K, y, flags=L.EPI_BIAS | L.EPI_SCALE | L.EPI_SCALE_TANH | L.EPI_RESID, bias=bias, resid=resid, scale=alpha)
    assert rel_l2(y, base * torch.tanh(alpha) + resid) < 1e-5
    y = torch.empty(M, N, device=DEV)
    ops.gemm(A, B, M, N, K, y, flags=L.EPI_BIAS | L.EPI_RESID, bias=bias, resid=resid)
    assert rel_l2(y, base + resid) < 1e-5


@pytest.mark.parametrize("D", [128, 512, 768, 1024])
def test_layernorm_fwd_bwd(D):
    ops, L = _ops()
    torch.manual_seed(2)
    rows = 333
    x = torch.randn(rows, D, device=DEV) * 2 + 0.5
    w, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    yb = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
    yf = torch.empty(rows, D, device=DEV)
    ops.layernorm_fwd(x, w, b, 1e-6, rows, D, y_bf16=yb, y_f32=yf)
    xr = x.clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = O.layer_norm(xr, wr, br, 1e-6)
    assert rel_l2(yf, ref) < 1e-5
    assert rel_l2(yb, ref) < 4e-3
    dy = torch.randn(rows, D, device=DEV)
    a1, a2 = torch.randn(rows, D, device=DEV), torch.randn(rows, D, device=DEV)
    ref.backward(dy.bfloat16().float())
    dx = torch.empty(rows, D, device=DEV)
    dxb = torch.empty(rows, D, device=DEV, dtype=torch.bfloat16)
    dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    ops.layernorm_bwd(dy.bfloat16(), x, w, 1e-6, rows, D, add1=a1, add2=a2, dx=dx, dx_bf16=dxb, dgamma=dg, dbeta=db)
    assert rel_l2(dx, xr.grad + a1 + a2) < 1e-4
    assert rel_l2(dxb, xr.grad + a1 + a2) < 4e-3
    assert rel_l2(dg, wr.grad) < 1e-4 and rel_l2(db, br.grad) < 1e-4
    # fp32 dy + strided rows (final norm on CLS rows)
    xs = torch.randn(7, 5, D, device=DEV)
    out = torch.empty(7, D, device=DEV)
    ops.layernorm_fwd(xs, w, b, 1e-6, 7, D, ldx=5 * D, y_f32=out)
    assert rel_l2(out, O.layer_norm(xs[:, 0], w, b, 1e-6)) < 1e-5


def _attn_case(mode, B, H, T, n, seed):
    """VarAttention (qkv GEMM + attention + proj GEMM) vs oracle.var_attention, forward and all gradients."""
    from lavila_b200.models.timesformer import VarAttention
    torch.manual_seed(seed)
    D = 64 * H
    N = 1 + T * n
    mod = VarAttention(D, num_heads=H, qkv_bias=True).to(DEV)
    with torch.no_grad():
        for p_ in mod.parameters():
            p_.normal_(0, 0.08)
    x = torch.randn(B, N, D, device=DEV, requires_grad=True)
    dims = {"f": T} if mode == "space" else {"n": n}
    y = mod(x, 'b (f n) d', '(b f) n d' if mode == "space" else '(b n) f d', dims)
    dy = torch.randn_like(y)
    y.backward(dy)
    p = {"a.qkv.weight": mod.qkv.weight.detach().clone().requires_grad_(True),
         "a.qkv.bias": mod.qkv.bias.detach().clone().requires_grad_(True),
         "a.proj.weight": mod.proj.weight.detach().clone().requires_grad_(True),
         "a.proj.bias": mod.proj.bias.detach().clone().requires_grad_(True)}
    xr = x.detach().clone().requires_grad_(True)
    ref = O.var_attention(xr, p, "a.", H, mode, T, n)
    ref.backward(dy)
    assert_close_bf16(y, ref, "%s attention out" % mode)
    assert_close_bf16(x.grad, xr.grad, "%s attention dx" % mode, rel=3e-2)
    assert_close_bf16(mod.qkv.weight.grad, p["a.qkv.weight"].grad, "%s d qkv.weight" % mode, rel=3e-2)
    assert_close_bf16(mod.qkv.bias.grad, p["a.qkv.bias"].grad, "%s d qkv.bias" % mode, rel=3e-2)
    assert_close_bf16(mod.proj.weight.grad, p["a.proj.weight"].grad, "%s d proj.weight" % mode, rel=3e-2)
    assert_close_bf16(mod.proj.bias.grad, p["a.proj.bias"].grad, "%s d proj.bias" % mode, rel=3e-2)


@pytest.mark.parametrize("mode,B,H,T,n", [("space", 2, 2, 4, 4), ("time", 2, 2, 4, 4), ("space", 1, 3, 2, 196),
                                          ("time", 1, 3, 16, 9), ("space", 2, 12, 16, 196), ("time", 2, 12, 16, 196),
                                          ("time", 2, 16, 8, 5), ("time", 3, 12, 16, 49)])
def test_var_attention(mode, B, H, T, n):
    _attn_case(mode, B, H, T, n, seed=3)


def test_mlp():
    from lavila_b200.models.timesformer import Mlp, QuickGELU
    torch.manual_seed(4)
    m = Mlp(256, 1024, act_layer=QuickGELU).to(DEV)
    x = torch.randn(5, 77, 256, device=DEV, requires_grad=True)
    y = m(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    xr = x.detach().clone().requires_grad_(True)
    w1, b1 = m.fc1.weight.detach().clone().requires_grad_(True), m.fc1.bias.detach().clone().requires_grad_(True)
    w2, b2 = m.fc2.weight.detach().clone().requires_grad_(True), m.fc2.bias.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.linear(O.quick_gelu(torch.nn.functional.linear(xr, w1, b1)), w2, b2)
    ref.backward(dy)
    assert_close_bf16(y, ref, "mlp out")
    assert_close_bf16(x.grad, xr.grad, "mlp dx", rel=3e-2)
    assert_close_bf16(m.fc1.weight.grad, w1.grad, "mlp dW1", rel=3e-2)
    assert_close_bf16(m.fc2.weight.grad, w2.grad, "mlp dW2", rel=3e-2)
    assert_close_bf16(m.fc1.bias.grad, b1.grad, "mlp db1", rel=3e-2)
    assert_close_bf16(m.fc2.bias.grad, b2.grad, "mlp db2", rel=3e-2)


def test_index_ops_bit_exact():
    ops, L = _ops()
    torch.manual_seed(5)
    B, Lc = 9, 77
    text = torch.randint(1, 49405, (B, Lc), device=DEV)
    for b in range(B):
        k = 3 + b
        text[b, k] = 49407
        text[b, k + 1:] = 0
    text[0, 5] = 49407  # duplicate maximum: first occurrence wins (torch.argmax)
    idx = torch.empty(B, device=DEV, dtype=torch.int32)
    ops.argmax_i64(text, idx, B, Lc)
    assert torch.equal(idx.long(), text.argmax(dim=-1))
    tok = torch.randn(49408, 128, device=DEV)
    pos = torch.randn(Lc, 128, device=DEV)
    x = torch.empty(B * Lc, 128, device=DEV)
    ops.text_embed(text, tok, pos, x, B * Lc, Lc, 128, 49408)
    assert torch.equal(x.view(B, Lc, 128), tok[text] + pos)     # gather + one fp32 add: bit exact


def test_clip_loss_single_rank():
    from lavila_b200.models.loss import CLIPLoss
    torch.manual_seed(6)
    N, E = 48, 256
    img = torch.nn.functional.normalize(torch.randn(N, E, device=DEV), dim=-1).requires_grad_(True)
    txt = torch.nn.functional.normalize(torch.randn(N, E, device=DEV) + 0.5 * img.detach(), dim=-1).requires_grad_(True)
    ls = torch.tensor(2.659, device=DEV, requires_grad=True)
    out = CLIPLoss()({"image_embed": img, "text_embed": txt, "logit_scale": ls.exp()})
    out["loss"].backward()
    ir, tr = img.detach().clone().requires_grad_(True), txt.detach().clone().requires_grad_(True)
    lr = ls.detach().clone().requires_grad_(True)
    ref = O.clip_loss(ir, tr, lr.exp())
    ref["loss"].backward()
    assert abs(float(out["loss"]) - float(ref["loss"])) < 1e-4
    assert float(out["clip_acc"]) == float(ref["clip_acc"])
    assert rel_l2(img.grad, ir.grad) < 1e-4 and rel_l2(txt.grad, tr.grad) < 1e-4
    assert abs(float(ls.grad) - float(lr.grad)) < 1e-4 * max(1.0, abs(float(lr.grad)))
