#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_narrator.py -m gpu -q -p no:cacheprovider ) > gpurun_out/r02_c13_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_c13_pytest.log | cut -c1-400
timeout 600 python tools/gpu_narrator_profile.py 10 24 2>&1 | grep -v "^=>\|Warning" | grep "us/step\|R=" | head -16 | cut -c1-200
timeout 600 python tools/gpu_narrator_profile.py 1 24 2>&1 | grep -v "^=>\|Warning" | grep "us/step\|R=" | head -14 | cut -c1-200
