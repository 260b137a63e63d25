"""Achieved HBM bandwidth of the memory-bound kernels at the BASELINE shapes (B=64, TSF-B 16f)."""
import sys
import torch
sys.path.insert(0, ".")
from lavila_b200 import ops

B, H, T, n = 64, 12, 16, 196
D = 64 * H
N = 1 + T * n
M = B * N
dev = "cuda"
torch.manual_seed(0)
u = M * D * 2 / 1e9  # GB of one bf16 [M, D] tensor


ONLY = sys.argv[1] if len(sys.argv) > 1 else None


def timeit(fn, iters=5):
    if ONLY is not None:
        return -1.0
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, gb):
    if ONLY is not None:
        return
    print("%-34s %8.3f ms  %7.2f GB algorithmic  %7.0f GB/s  (%.2f of 6582)" % (name, ms, gb, gb / ms * 1e3, gb / ms * 1e3 / 6582))


x = torch.randn(M, D, device=dev)
w, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
yb = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
report("ln_fwd (f32 -> bf16)", timeit(lambda: ops.layernorm_fwd(x, w, b, 1e-6, M, D, y_bf16=yb)), 3 * u)
dyb = torch.randn(M, D, device=dev).bfloat16()
a1, a2 = torch.randn(M, D, device=dev), torch.randn(M, D, device=dev)
dx = torch.empty(M, D, device=dev)
dxb = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
report("ln_bwd (no adds, f32+bf16 out)", timeit(lambda: ops.layernorm_bwd(dyb, x, w, 1e-6, M, D, dx=dx, dx_bf16=dxb, dgamma=dg, dbeta=db)), 6 * u)
report("ln_bwd (1 add)", timeit(lambda: ops.layernorm_bwd(dyb, x, w, 1e-6, M, D, add1=a1, dx=dx, dx_bf16=dxb, dgamma=dg, dbeta=db)), 8 * u)
if ONLY == "ln_bwd":
    ops.layernorm_bwd(dyb, x, w, 1e-6, M, D, add1=a1, dx=dx, dx_bf16=dxb, dgamma=dg, dbeta=db)
    torch.cuda.synchronize()
    sys.exit(0)
report("ln_bwd (2 adds)", timeit(lambda: ops.layernorm_bwd(dyb, x, w, 1e-6, M, D, add1=a1, add2=a2, dx=dx, dx_bf16=dxb, dgamma=dg, dbeta=db)), 10 * u)
del a1, a2, dx, dxb, x
qkv = torch.randn(M, 3 * D, device=dev).bfloat16()
out = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(M, H, device=dev)
report("time attn fwd", timeit(lambda: ops.group_attn_fwd(qkv, out, lse, 1, B, H, T=T, n=n)), 4 * u)
report("space attn fwd (tcgen05)", timeit(lambda: ops.group_attn_fwd(qkv, out, lse, 0, B, H, T=T, n=n)), 4 * u)
report("cls attn fwd", timeit(lambda: ops.cls_attn_fwd(qkv, out, lse, B, H, N)), 2 * u)
ops.group_attn_fwd(qkv, out, lse, 1, B, H, T=T, n=n)
ops.cls_attn_fwd(qkv, out, lse, B, H, N)
dout = torch.randn(M, D, device=dev).bfloat16()
dqkv = torch.zeros(M, 3 * D, device=dev, dtype=torch.bfloat16)
dcls = torch.zeros(B, H, 2, 64, device=dev)
report("cls attn bwd (accumulate)", timeit(lambda: ops.cls_attn_bwd(qkv, out, dout, lse, dqkv, dcls, B, H, N, accumulate=True)), 6 * u)
if ONLY == "time_bwd":
    ops.group_attn_bwd(qkv, out, lse, dout, dqkv, dcls, 1, 1, B, H, T=T, n=n)
    torch.cuda.synchronize()
    sys.exit(0)
report("time attn bwd (accumulate kv)", timeit(lambda: ops.group_attn_bwd(qkv, out, lse, dout, dqkv, dcls, 1, 1, B, H, T=T, n=n)), 10 * u)
report("time attn bwd (no accumulate)", timeit(lambda: ops.group_attn_bwd(qkv, out, lse, dout, dqkv, dcls, 0, 1, B, H, T=T, n=n)), 8 * u)
report("space attn bwd (tcgen05, acc)", timeit(lambda: ops.group_attn_bwd(qkv, out, lse, dout, dqkv, dcls, 1, 0, B, H, T=T, n=n)), 10 * u)
bias = torch.zeros(3 * D, device=dev)
report("colsum [M, 3D]", timeit(lambda: ops.colsum_bf16(dqkv, M, 3 * D, bias)), 3 * u)
report("colsum [M, D]", timeit(lambda: ops.colsum_bf16(dout, M, D, bias)), 1 * u)
