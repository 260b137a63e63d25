"""Decode-sized GEMM (lv_gemm_skinny_bf16) timing per shape; run with LAVILA_B200_SKINNY_TC=0 / 1 to compare the mma.sync
kernel with the tcgen05 swap-AB route for more than 64 rows."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lavila_b200 import ops, _lib as L  # noqa: E402

dev = "cuda"
print("LAVILA_B200_SKINNY_TC =", os.environ.get("LAVILA_B200_SKINNY_TC", "(default 1)"))
for M in (128, 320):
    for (N, K) in ((4800, 1600), (1600, 1600), (6400, 1600), (1600, 6400)):
        A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        Ws = [(torch.randn(K, N, device=dev) * 0.05).bfloat16() for _ in range(8)]     # 8 weights: no L2 reuse across calls
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for w in Ws[:2]:
            ops.gemm_skinny(A, w, M, N, K, out, flags=L.EPI_BIAS, bias=bias)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(40):
            ops.gemm_skinny(A, Ws[it % 8], M, N, K, out, flags=L.EPI_BIAS, bias=bias)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 40 * 1e3
        print("M=%d N=%d K=%d  %.1f us  %.0f TFLOP/s  W stream %.0f GB/s" % (M, N, K, us, 2.0 * M * N * K / us / 1e6, 2.0 * N * K / us / 1e3))
