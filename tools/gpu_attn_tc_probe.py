"""Compare the tcgen05 space attention kernels with the mma.sync ones (already validated against the oracle)."""
import sys
import torch
sys.path.insert(0, ".")
from lavila_b200 import ops

which = sys.argv[1]
B, H, T, n = (int(x) for x in sys.argv[2:6]) if len(sys.argv) > 5 else (2, 12, 16, 196)
D = 64 * H
N = 1 + T * n
M = B * N
torch.manual_seed(0)
dev = "cuda"
qkv = (torch.randn(M, 3 * D, device=dev) * 1.0).bfloat16()
qkv[:, :D] *= 2.0


def run_fwd(tc):
    ops.USE_TC_ATTN_FWD = tc
    out = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(M, H, device=dev)
    ops.group_attn_fwd(qkv, out, lse, 0, B, H, T=T, n=n)
    ops.cls_attn_fwd(qkv, out, lse, B, H, N)
    torch.cuda.synchronize()
    return out, lse


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


o_ref, l_ref = run_fwd(False)
if which == "fwd":
    o, l = run_fwd(True)
    print("fwd out rel %.3e  lse rel %.3e  max|lse diff| %.3e" % (rel(o, o_ref), rel(l, l_ref), float((l - l_ref).abs().max())))
    for _ in range(3):
        run_fwd(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for tc in (True, False):
        ops.USE_TC_ATTN_FWD = tc
        out = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
        lse = torch.zeros(M, H, device=dev)
        e0.record()
        for _ in range(5):
            ops.group_attn_fwd(qkv, out, lse, 0, B, H, T=T, n=n)
        e1.record()
        torch.cuda.synchronize()
        print("fwd tc=%s %.3f ms" % (tc, e0.elapsed_time(e1) / 5))
    print("RESULT fwd", "PASS" if rel(o, o_ref) < 1e-2 and rel(l, l_ref) < 1e-3 else "FAIL")
else:
    dout = torch.randn(M, D, device=dev).bfloat16()

    def run_bwd(tc, group_first=True):
        ops.USE_TC_ATTN_BWD = tc
        dqkv = torch.full((M, 3 * D), 7.0, device=dev, dtype=torch.bfloat16)   # every element must be overwritten
        dcls = torch.zeros(B, H, 2, 64, device=dev)
        if group_first:   # the engine's order: plain writes, then the accumulating CLS-query pass
            ops.group_attn_bwd(qkv, o_ref, l_ref, dout, dqkv, dcls, 0, 0, B, H, T=T, n=n)
            ops.cls_attn_bwd(qkv, o_ref, dout, l_ref, dqkv, dcls, B, H, N, accumulate=True)
        else:
            ops.cls_attn_bwd(qkv, o_ref, dout, l_ref, dqkv, dcls, B, H, N)
            ops.group_attn_bwd(qkv, o_ref, l_ref, dout, dqkv, dcls, 1, 0, B, H, T=T, n=n)
        ops.cls_kv_finalize(dcls, dqkv, B, H, N)
        torch.cuda.synchronize()
        return dqkv
    g_ref = run_bwd(False)
    g2 = run_bwd(True, group_first=False)
    print("bwd (cls first, accumulate path) rel %.3e" % rel(g2, g_ref))
    g = run_bwd(True)
    for nm, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        print("bwd %s rel %.3e   cls-row rel %.3e" % (nm, rel(g[:, sl], g_ref[:, sl]), rel(g[::N, sl], g_ref[::N, sl])))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for tc in (True, False):
        ops.USE_TC_ATTN_BWD = tc
        dqkv = torch.zeros(M, 3 * D, device=dev, dtype=torch.bfloat16)
        dcls = torch.zeros(B, H, 2, 64, device=dev)
        e0.record()
        for _ in range(5):
            ops.group_attn_bwd(qkv, o_ref, l_ref, dout, dqkv, dcls, 0, 0, B, H, T=T, n=n)
        e1.record()
        torch.cuda.synchronize()
        print("bwd tc=%s %.3f ms" % (tc, e0.elapsed_time(e1) / 5))
    print("RESULT bwd", "PASS" if rel(g, g_ref) < 2e-2 else "FAIL")
