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
                                          ("time", 2, 16, 8, 5), ("time", 3, 12, 16, 49),
                                          # TSF-L/14 geometries: 256 / 576 patches per frame (key-tiled kernels), 32 frames
                                          ("space", 1, 2, 2, 256), ("space", 2, 3, 3, 576), ("space", 1, 2, 2, 209),
                                          ("time", 1, 2, 32, 5), ("time", 1, 1, 250, 2)])
def test_var_attention(mode, B, H, T, n):
    _attn_case(mode, B, H, T, n, seed=3)


@pytest.mark.parametrize("mode,B,H,T,n", [(0, 1, 1, 1, 320), (0, 2, 2, 3, 257), (0, 1, 16, 2, 576), (1, 1, 2, 300, 3)])
def test_tiled_group_attention_direct(mode, B, H, T, n):
    """lv_group_attn_fwd / _bwd on groups of more than 208 keys (attention_big.cu) against fp32 torch math on the same
    bf16 qkv: out, lse, dq / dk / dv of the patch rows and the CLS key/value gradient accumulator."""
    ops, _ = _ops()
    torch.manual_seed(11)
    D, N = 64 * H, 1 + T * n
    qkv = (torch.randn(B * N, 3 * D, device=DEV) * 1.5).to(torch.bfloat16)
    dout = torch.randn(B * N, D, device=DEV).to(torch.bfloat16)
    out = torch.zeros(B * N, D, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(B * N, H, device=DEV)
    ops.group_attn_fwd(qkv, out, lse, mode, B, H, T=T, n=n)
    # reference: [B, H, groups, Lq, 64] queries; keys = group rows + CLS
    x = qkv.float().view(B, N, 3, H, 64).requires_grad_(True)
    q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]                      # [B, N, H, 64]

    def grp(t):   # patch rows -> [B, H, G, Lq, 64]
        t = t[:, 1:].reshape(B, T, n, H, 64)
        return t.permute(0, 3, 1, 2, 4) if mode == 0 else t.permute(0, 3, 2, 1, 4)
    G = T if mode == 0 else n
    qg, kg, vg = grp(q), grp(k), grp(v)
    kc = k[:, :1].permute(0, 2, 1, 3)[:, :, None].expand(B, H, G, 1, 64)
    vc = v[:, :1].permute(0, 2, 1, 3)[:, :, None].expand(B, H, G, 1, 64)
    ka, va = torch.cat([kg, kc], 3), torch.cat([vg, vc], 3)
    sc = torch.einsum("bhgqd,bhgkd->bhgqk", qg, ka) * 0.125
    ref_lse = torch.logsumexp(sc, -1)
    ref_o = torch.einsum("bhgqk,bhgkd->bhgqd", sc.softmax(-1), va)

    def ungrp(t):  # [B, H, G, Lq, 64] -> [B, T*n, H, 64]
        t = t.permute(0, 2, 3, 1, 4) if mode == 0 else t.permute(0, 3, 2, 1, 4)
        return t.reshape(B, T * n, H, 64)
    got_o = out.view(B, N, H, 64)[:, 1:]
    assert_close_bf16(got_o, ungrp(ref_o), "tiled attention out")
    lg = lse.view(B, N, H)[:, 1:]
    lr = (ref_lse.permute(0, 2, 3, 1) if mode == 0 else ref_lse.permute(0, 3, 2, 1)).reshape(B, T * n, H)
    lse_err = float((lg - lr.detach()).abs().max())
    assert lse_err < 2e-2, "lse max abs diff %.3e" % lse_err
    # backward: feed the kernel's own (bf16) forward output, as the engine does
    do = dout.float().view(B, N, H, 64)[:, 1:]
    (ungrp(ref_o) * do).sum().backward()
    dqkv = torch.zeros(B * N, 3 * D, device=DEV, dtype=torch.bfloat16)
    dcls = torch.zeros(B, H, 2, 64, device=DEV)
    ops.group_attn_bwd(qkv, out, lse, dout, dqkv, dcls, 0, mode, B, H, T=T, n=n)
    gd = dqkv.float().view(B, N, 3, H, 64)
    for i, nm in enumerate(["dq", "dk", "dv"]):
        assert_close_bf16(gd[:, 1:, i], x.grad[:, 1:, i], "tiled attention " + nm, rel=3e-2)
    assert_close_bf16(dcls[:, :, 0], x.grad[:, 0, 1], "tiled attention d cls key", rel=3e-2)
    assert_close_bf16(dcls[:, :, 1], x.grad[:, 0, 2], "tiled attention d cls value", rel=3e-2)
    assert float(gd[:, 0].abs().max()) == 0.0, "the group kernels must not touch the CLS row of dqkv"


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


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_ssl_clip_loss_matches_reference_golden(idx):
    """SSLCLIPLoss (lavila/models/loss.py:121-217) through the fused kernels against golden vectors of the UNMODIFIED
    reference (tests/golden/make_golden_ssl.py): loss, the three accuracies (NaN for an empty class, as the reference),
    the counts (bit exact) and the gradients w.r.t. embeddings, logit_scale and logit_scale_pseudo.  fp32 -> 1e-4."""
    import os
    from lavila_b200.models.loss import SSLCLIPLoss
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ssl_loss_small.pt"), weights_only=False)
    c = G["world1"][idx]
    crit = SSLCLIPLoss(scale_init=G["scale_init"]).to(DEV)
    img = c["image"].to(DEV).requires_grad_(True)
    txt = c["text"].to(DEV).requires_grad_(True)
    s = torch.tensor(G["scale"], device=DEV, requires_grad=True)
    out = crit({"image_embed": img, "text_embed": txt, "logit_scale": s}, c["gt"].to(DEV))
    gi, gt, gs, gp = torch.autograd.grad(out["loss"], (img, txt, s, crit.logit_scale_pseudo))
    assert abs(float(out["loss"]) - float(c["loss"])) < 1e-4
    for k in ("clip_acc", "clip_acc_gt", "clip_acc_pseudo"):
        a, b = float(out[k]), float(c[k])
        assert (a != a and b != b) or abs(a - b) < 1e-3, (k, a, b)
    assert int(out["num_gt"]) == int(c["num_gt"]) and int(out["num_pseudo"]) == int(c["num_pseudo"])
    assert out["num_gt"].shape == c["num_gt"].shape and out["num_gt"].dtype == c["num_gt"].dtype
    assert rel_l2(gi, c["grad_image"]) < 1e-4 and rel_l2(gt, c["grad_text"]) < 1e-4
    assert abs(float(gs) - float(c["grad_scale"])) < 1e-5 + 1e-4 * abs(float(c["grad_scale"]))
    assert abs(float(gp) - float(c["grad_scale_pseudo"])) < 1e-5 + 1e-4 * abs(float(c["grad_scale_pseudo"]))


def test_ssl_clip_loss_two_rank_rows():
    """The 2-rank golden (reference SSLCLIPLoss(use_vissl) over gloo): the kernels evaluated on the gathered batch give
    every rank's loss, and the local-row gradient slices times W equal the reference's GatherLayer gradients."""
    import os
    ops, _ = _ops()
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ssl_loss_small.pt"), weights_only=False)
    r = G["world2"]
    W, B = len(r), r[0]["image"].shape[0]
    img = torch.cat([x["image"] for x in r]).to(DEV).contiguous()
    txt = torch.cat([x["text"] for x in r]).to(DEV).contiguous()
    gt = torch.cat([x["gt"] for x in r]).to(DEV).to(torch.int32).contiguous()
    Ng, E = img.shape
    s = torch.tensor([G["scale"]], device=DEV)
    sp = torch.tensor([1.0 / G["scale_init"]], device=DEV)
    lse_i, lse_t = torch.empty(Ng, device=DEV), torch.empty(Ng, device=DEV)
    partial, result = torch.empty(2 * Ng, device=DEV), torch.empty(6, device=DEV)
    counter = torch.zeros(1, device=DEV, dtype=torch.int32)
    ops.ssl_clip_loss_fwd(img, txt, s, sp, gt, Ng, E, lse_i, lse_t, partial, counter, result)
    g = torch.ones(1, device=DEV)
    for k in range(W):
        assert abs(float(result[0]) - float(r[k]["loss"])) < 1e-4
        d_i, d_t, d_s = torch.empty(B, E, device=DEV), torch.empty(B, E, device=DEV), torch.zeros(2, device=DEV)
        ops.ssl_clip_loss_bwd(img, txt, s, sp, gt, lse_i, lse_t, g, float(W), 1.0, Ng, E, k * B, B, d_i, d_t, d_s)
        assert rel_l2(d_i, r[k]["grad_image"]) < 1e-4 and rel_l2(d_t, r[k]["grad_text"]) < 1e-4
    d_i, d_t, d_s = torch.empty(Ng, E, device=DEV), torch.empty(Ng, E, device=DEV), torch.zeros(2, device=DEV)
    ops.ssl_clip_loss_bwd(img, txt, s, sp, gt, lse_i, lse_t, g, float(W), 1.0, Ng, E, 0, Ng, d_i, d_t, d_s)
    assert abs(float(d_s[0]) - float(r[0]["grad_scale"])) < 1e-5 + 1e-4 * abs(float(r[0]["grad_scale"]))
    # the reference's gradient is w.r.t. the log-parameter: d/d log(s_p) = s_p * d/d s_p
    assert abs(float(d_s[1] * sp[0]) - float(r[0]["grad_scale_pseudo"])) < 1e-5 + 1e-4 * abs(float(r[0]["grad_scale_pseudo"]))
