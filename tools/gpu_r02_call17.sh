#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gpu_cls_fusion_probe.py 2 2 4 196 2>&1 | tail -10
timeout 120 python tools/gpu_cls_fusion_probe.py 64 12 16 196 2>&1 | tail -10
timeout 120 python tools/gpu_attn_tc_probe.py bwd 64 12 16 196 2>&1 | tail -7
( timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_realsize_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider ) > gpurun_out/r02_c17_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_c17_pytest.log | cut -c1-300
