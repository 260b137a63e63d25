#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do ( timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider ) > gpurun_out/r02_c18_pytest_$i.log 2>&1; echo "pytest run $i rc=$?"; tail -1 gpurun_out/r02_c18_pytest_$i.log | cut -c1-200; done
echo "== memcheck: gemm epilogues"; timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu_gemm_probe.py epi 1000 768 512 2>&1 | tail -4
echo "== memcheck: cls fusion"; timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu_cls_fusion_probe.py 2 2 2 196 2>&1 | grep -v " ms$" | tail -5
echo "== memcheck: flash tc + sampling + narrator tests"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_narrator.py -m gpu -q -x -p no:cacheprovider -k "flash or top_p or skinny" 2>&1 | tail -4
echo "== racecheck: flash tc"; timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_narrator.py -m gpu -q -x -p no:cacheprovider -k "flash_attention and (200 or 1025 or 257)" 2>&1 | tail -4
