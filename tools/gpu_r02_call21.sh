#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gpu_cls_fusion_probe.py 64 12 16 196 1 2>&1 | tail -6
echo "== rows ON"; timeout 120 python tools/gpu_gemm2_epi_probe.py all 20
echo "== rows OFF"; LAVILA_B200_GEMM_ROWS_EPI=0 timeout 120 python tools/gpu_gemm2_epi_probe.py proj 20; LAVILA_B200_GEMM_ROWS_EPI=0 timeout 120 python tools/gpu_gemm2_epi_probe.py fc2 20
( timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider ) > gpurun_out/r02_c21_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_c21_pytest.log | cut -c1-300
for v in 1 0; do
LAVILA_B200_GEMM_ROWS_EPI=$v timeout 400 python bench.py --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline --no-narrator --no-e2e > gpurun_out/r02_c21_bench_$v.json 2> gpurun_out/r02_c21_bench.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_c21_bench_$v.json').read())
print("rows=$v", {k:d[k] for k in ('value','ms_per_step')}, d['block_roofline']['frac'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['clocks']['sm_mhz'])
for c in d['roofline']['by_class']:
    if c['flags'] in (33,): print("   ", c)
PY
done
