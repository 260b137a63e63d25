"""Randomized sweeps of the CUDA entry points against their torch test doubles (tests/ops_doubles.py): the same call, with the
same arguments, on the device (kernel) and on CPU copies (double); every output buffer is compared.  The doubles themselves are
pinned to the unmodified reference through the host-logic CPU tests, so this closes the loop per entry point over many more shapes,
strides and flag combinations than the hand-written parity tests.

Fixed seeds; part of `-m gpu` since round 2 (also selectable alone with `-m gpu_fuzz`).
"""
import random

import pytest
import torch

from tests import ops_doubles as D
from tests.util import cosine, rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.gpu_fuzz]
DEV = "cuda"
BF16, F32 = torch.bfloat16, torch.float32


def _cmp(got, want, what, rel=2e-2):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert torch.isfinite(got).all(), what
    if float(want.norm()) < 1e-12:
        assert float(got.norm()) < 1e-6, what
        return
    r = rel_l2(got, want)
    assert r < rel and cosine(got, want) > 0.999, "%s: rel_l2 %.3e" % (what, r)


def _both(fn_name, dev_args, dev_kwargs):
    """Call the kernel wrapper on the device tensors and the double on CPU clones; returns the list of (dev, cpu) tensor pairs."""
    from lavila_b200 import ops
    pairs = []

    def clone(x):
        if torch.is_tensor(x):
            c = x.detach().cpu().clone()
            pairs.append((x, c))
            return c
        return x
    cpu_args = [clone(a) for a in dev_args]
    cpu_kwargs = {k: clone(v) for k, v in dev_kwargs.items()}
    getattr(ops, fn_name)(*dev_args, **dev_kwargs)
    getattr(D, fn_name)(*cpu_args, **cpu_kwargs)
    torch.cuda.synchronize()
    return pairs


@pytest.mark.parametrize("seed", range(12))
def test_gemm_fuzz(seed):
    from lavila_b200 import _lib as L
    rng = random.Random(seed)
    M = rng.choice([1, 7, 64, 130, 257, 1000])
    N = rng.choice([64, 128, 256, 320, 768])
    K = rng.choice([64, 128, 192, 768])
    a_mn, b_mn = rng.choice([(0, 0), (0, 1), (1, 1)])
    if a_mn and M % 8:
        M = (M + 7) // 8 * 8          # an MN-major A has row pitch M: TMA needs a 16-byte multiple
    torch.manual_seed(seed)
    A = (torch.randn((K, M) if a_mn else (M, K), device=DEV) * 0.5).to(BF16)
    Bm = (torch.randn((K, N) if b_mn else (N, K), device=DEV) * 0.1).to(BF16)
    bias, resid, gate = torch.randn(N, device=DEV), torch.randn(M, N, device=DEV), torch.tensor(0.3, device=DEV)
    choice = rng.choice(["plain", "bias", "bias_resid", "gate_resid", "quickgelu", "dquickgelu", "gelu_tanh", "sqrelu", "atomic"])
    kw, out_dtype = dict(a_mn=a_mn, b_mn=b_mn), BF16
    if choice == "bias":
        kw.update(flags=L.EPI_BIAS, bias=bias)
    elif choice == "bias_resid":
        kw.update(flags=L.EPI_BIAS | L.EPI_RESID, bias=bias, resid=resid)
        out_dtype = F32
    elif choice == "gate_resid":
        kw.update(flags=L.EPI_BIAS | L.EPI_RESID | L.EPI_SCALE | L.EPI_SCALE_TANH, bias=bias, resid=resid, scale=gate)
        out_dtype = F32
    elif choice == "quickgelu":
        kw.update(flags=L.EPI_BIAS | L.EPI_QUICKGELU, bias=bias, out2=torch.zeros(M, N, device=DEV, dtype=BF16))
    elif choice == "dquickgelu":
        kw.update(flags=L.EPI_DQUICKGELU, aux=torch.randn(M, N, device=DEV).to(BF16))
    elif choice == "gelu_tanh":
        kw.update(flags=L.EPI_BIAS | L.EPI_GELU_TANH, bias=bias)
    elif choice == "sqrelu":
        kw.update(flags=L.EPI_BIAS | L.EPI_SQRELU, bias=bias)
    elif choice == "atomic":
        kw.update(flags=L.EPI_ATOMIC, k_splits=rng.choice([1, 2, 3]) if K >= 192 else 1)
        out_dtype = F32
    out = torch.zeros(M, N, device=DEV, dtype=out_dtype)
    pairs = _both("gemm", [A, Bm, M, N, K, out], kw)
    for dev_t, cpu_t in pairs:
        if dev_t is out or dev_t is kw.get("out2"):
            _cmp(dev_t, cpu_t, "gemm %s M=%d N=%d K=%d a_mn=%d b_mn=%d" % (choice, M, N, K, a_mn, b_mn))


@pytest.mark.parametrize("seed", range(8))
def test_layernorm_fuzz(seed):
    rng = random.Random(100 + seed)
    D_ = rng.choice([128, 256, 768, 1024])
    rows = rng.choice([1, 5, 33, 1000])
    torch.manual_seed(seed)
    x = torch.randn(rows, D_, device=DEV) * 2 + 0.5
    g, b = torch.randn(D_, device=DEV), torch.randn(D_, device=DEV)
    y = torch.zeros(rows, D_, device=DEV, dtype=BF16)
    for dev_t, cpu_t in _both("layernorm_fwd", [x, g, b, 1e-6, rows, D_], dict(y_bf16=y)):
        if dev_t is y:
            _cmp(dev_t, cpu_t, "ln fwd")
    dy = torch.randn(rows, D_, device=DEV).to(BF16)
    add = torch.randn(rows, D_, device=DEV)
    dx, dxb = torch.zeros(rows, D_, device=DEV), torch.zeros(rows, D_, device=DEV, dtype=BF16)
    dg, db = torch.zeros(D_, device=DEV), torch.zeros(D_, device=DEV)
    for dev_t, cpu_t in _both("layernorm_bwd", [dy, x, g, 1e-6, rows, D_], dict(add1=add, dx=dx, dx_bf16=dxb, dgamma=dg, dbeta=db)):
        if any(dev_t is t for t in (dx, dxb, dg, db)):
            _cmp(dev_t, cpu_t, "ln bwd", rel=3e-2)


@pytest.mark.parametrize("seed", range(10))
def test_group_and_cls_attention_fuzz(seed):
    rng = random.Random(200 + seed)
    mode = rng.choice([0, 1, 2])
    B, H = rng.choice([1, 2, 3]), rng.choice([1, 2, 12])
    T, n, Lctx = rng.choice([1, 4, 16, 32]), rng.choice([1, 9, 49, 196, 256]), rng.choice([5, 16, 77])
    if mode == 2:
        rows, kw = B * Lctx, dict(Lctx=Lctx)
    else:
        rows, kw = B * (1 + T * n), dict(T=T, n=n)
    D_ = 64 * H
    torch.manual_seed(seed)
    qkv = (torch.randn(rows, 3 * D_, device=DEV) * 1.2).to(BF16)
    out = torch.zeros(rows, D_, device=DEV, dtype=BF16)
    lse = torch.zeros(rows, H, device=DEV)
    for dev_t, cpu_t in _both("group_attn_fwd", [qkv, out, lse, mode, B, H], kw):
        if dev_t is out or dev_t is lse:
            _cmp(dev_t, cpu_t, "group fwd mode %d B%d H%d T%d n%d L%d" % (mode, B, H, T, n, Lctx))
    dout = torch.randn(rows, D_, device=DEV).to(BF16)
    dqkv = torch.zeros(rows, 3 * D_, device=DEV, dtype=BF16)
    dcls = torch.zeros(B, H, 2, 64, device=DEV) if mode != 2 else None
    for dev_t, cpu_t in _both("group_attn_bwd", [qkv, out, lse, dout, dqkv, dcls, 0, mode, B, H], kw):
        if dev_t is dqkv or (dcls is not None and dev_t is dcls):
            _cmp(dev_t, cpu_t, "group bwd mode %d B%d H%d T%d n%d L%d" % (mode, B, H, T, n, Lctx), rel=3e-2)
    if mode != 2:
        N = 1 + T * n
        for dev_t, cpu_t in _both("cls_attn_fwd", [qkv, out, lse, B, H, N], {}):
            if dev_t is out or dev_t is lse:
                _cmp(dev_t, cpu_t, "cls fwd")
        for dev_t, cpu_t in _both("cls_attn_bwd", [qkv, out, dout, lse, dqkv, dcls, B, H, N], dict(accumulate=True)):
            if dev_t is dqkv or dev_t is dcls:
                _cmp(dev_t, cpu_t, "cls bwd", rel=3e-2)


@pytest.mark.parametrize("seed", range(8))
def test_flash_and_skinny_fuzz(seed):
    from lavila_b200 import _lib as L
    rng = random.Random(300 + seed)
    B, H, Lq, Lk = rng.choice([1, 3]), rng.choice([1, 12, 25]), rng.choice([1, 5, 77]), rng.choice([1, 40, 256, 300])
    causal = rng.choice([False, True]) and Lq <= Lk
    torch.manual_seed(seed)
    q = torch.randn(B * Lq, H * 64, device=DEV).to(BF16)
    k = torch.randn(B * Lk, H * 64, device=DEV).to(BF16)
    v = torch.randn(B * Lk, H * 64, device=DEV).to(BF16)
    out = torch.zeros(B * Lq, H * 64, device=DEV, dtype=BF16)
    pairs = _both("flash_attn_fwd", [q, k, v, out, B, H, Lq, Lk],
                  dict(q_rows=Lq, kv_rows=Lk, ld_q=H * 64, ld_kv=H * 64, ld_out=H * 64, causal=causal))
    for dev_t, cpu_t in pairs:
        if dev_t is out:
            _cmp(dev_t, cpu_t, "flash B%d H%d Lq%d Lk%d causal=%s" % (B, H, Lq, Lk, causal))
    M, N, K = rng.choice([1, 32, 100, 320]), rng.choice([64, 768, 1600]), rng.choice([64, 768, 1600])
    A = (torch.randn(M, K, device=DEV) * 0.5).to(BF16)
    W = (torch.randn(K, N, device=DEV) * 0.05).to(BF16)
    bias = torch.randn(N, device=DEV)
    o2 = torch.zeros(M, N, device=DEV, dtype=BF16)
    for dev_t, cpu_t in _both("gemm_skinny", [A, W, M, N, K, o2], dict(flags=L.EPI_BIAS | L.EPI_GELU_TANH, bias=bias)):
        if dev_t is o2:
            _cmp(dev_t, cpu_t, "skinny M%d N%d K%d" % (M, N, K), rel=1e-2)
