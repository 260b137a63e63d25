#!/bin/bash
mkdir -p gpurun_out
for v in 1 0 1 0; do
LAVILA_B200_CLS_FUSION=$v timeout 400 python bench.py --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline --no-narrator --no-e2e > gpurun_out/r02_c15_bench_$v.json 2> gpurun_out/r02_c15_bench.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_c15_bench_$v.json').read())
print("fusion=$v", {k:d[k] for k in ('value','ms_per_step')}, d['block_roofline']['frac'], d['block_roofline']['ms'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['clocks']['sm_mhz'])
PY
done
