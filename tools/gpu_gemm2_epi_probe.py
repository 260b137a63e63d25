"""Standalone timing of the hot GEMM classes through ops.gemm (2-CTA kernel): TF/s per class; used under ncu for the epilogue study."""
import sys
import torch
sys.path.insert(0, ".")
from lavila_b200 import ops, _lib as L

dev = "cuda"
M = 200768
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cases = {
    "qkv": (M, 2304, 768, 0, L.EPI_BIAS),
    "fc1": (M, 3072, 768, 0, L.EPI_BIAS | L.EPI_QUICKGELU),
    "dgelu": (M, 3072, 768, 1, L.EPI_DQUICKGELU),
    "dgrad": (M, 768, 2304, 1, 0),
    "proj": (M, 768, 768, 0, L.EPI_BIAS | L.EPI_RESID),
    "fc2": (M, 768, 3072, 0, L.EPI_BIAS | L.EPI_RESID),
}
torch.manual_seed(0)
for name, (m, n, k, b_mn, fl) in cases.items():
    if which != "all" and which != name:
        continue
    A = (torch.randn(m, k, device=dev) * 0.1).bfloat16()
    B = (torch.randn((k, n) if b_mn else (n, k), device=dev) * 0.1).bfloat16()
    f32 = bool(fl & L.EPI_RESID)
    out = torch.empty(m, n, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    kw = {}
    if fl & L.EPI_BIAS:
        kw["bias"] = torch.randn(n, device=dev)
    if fl & L.EPI_RESID:
        kw["resid"] = torch.randn(m, n, device=dev)
    if fl & L.EPI_QUICKGELU:
        kw["out2"] = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    if fl & L.EPI_DQUICKGELU:
        kw["aux"] = torch.randn(m, n, device=dev).bfloat16()
    for _ in range(2):
        ops.gemm(A, B, m, n, k, out, b_mn=b_mn, flags=fl, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.gemm(A, B, m, n, k, out, b_mn=b_mn, flags=fl, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-6s %dx%dx%d flags %d: %.3f ms %.1f TF/s" % (name, m, n, k, fl, ms, 2.0 * m * n * k / ms / 1e9))
