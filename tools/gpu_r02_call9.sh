#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_narrator.py tests/test_gpu_ops.py tests/test_gpu_realsize_parity.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider ) > gpurun_out/r02_c9_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_c9_pytest.log | cut -c1-300
echo "== flash tc ON"; timeout 120 python tools/gpu_flash_tc_probe.py 2>&1 | tail -6
echo "== flash tc OFF"; LAVILA_B200_FLASH_TC=0 timeout 120 python tools/gpu_flash_tc_probe.py 2>&1 | tail -6
