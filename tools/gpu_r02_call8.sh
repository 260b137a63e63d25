#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gpu_gemm_probe.py epi 1000 768 512 2>&1 | tail -7
timeout 120 python tools/gpu_gemm_probe.py epi 4096 324 768 2>&1 | tail -7
echo "== rows ON"; timeout 120 python tools/gpu_gemm2_epi_probe.py all 20
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fuzz_vs_doubles.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
