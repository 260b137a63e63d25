#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_narrator.py tests/test_gpu_fuzz_vs_doubles.py tests/test_gpu_realsize_parity.py -m gpu -q -x -p no:cacheprovider ) > gpurun_out/r02_c16_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_c16_pytest.log | cut -c1-300
timeout 300 python tools/bench_narrator.py --encoder large --batch 32 --returns 1,10 2>&1 | grep impl | cut -c1-330
timeout 600 python tools/gpu_narrator_profile.py 10 24 2>&1 | grep -v "^=>\|Warning" | grep "us/step\|R=" | head -9 | cut -c1-200
