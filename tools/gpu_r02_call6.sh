#!/bin/bash
mkdir -p gpurun_out
for v in "epi 1000 768 512" "epi 4096 320 768" "nt 4096 2304 768" "nn 1000 768 768"; do echo "=== $v"; timeout 120 python tools/gpu_gemm_probe.py $v 2>&1 | tail -8; done
( time timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider ) > gpurun_out/r02_c6_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_c6_pytest.log | cut -c1-300
echo "=== perf rows epilogue ON"; timeout 200 python tools/gpu_gemm_probe.py perf 2>&1 | tail -13
echo "=== perf rows epilogue OFF"; LAVILA_B200_GEMM_ROWS_EPI=0 timeout 200 python tools/gpu_gemm_probe.py perf 2>&1 | tail -13
timeout 400 python bench.py --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/r02_c6_bench.json 2> gpurun_out/r02_c6_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c6_bench.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['block_roofline']['frac'], d['block_roofline']['ms'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['clocks'])
for c in d['roofline']['by_class']: print(c)
PY
