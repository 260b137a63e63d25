#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_realsize_parity.py tests/test_gpu_fuzz_vs_doubles.py -m gpu -q -p no:cacheprovider ) > gpurun_out/r02_c2_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r02_c2_pytest.log | cut -c1-300
timeout 120 python tools/gpu_attn_bwd_stamps.py > gpurun_out/r02_c2_stamps.log 2>&1; cat gpurun_out/r02_c2_stamps.log
