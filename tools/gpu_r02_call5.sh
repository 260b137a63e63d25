#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider ) > gpurun_out/r02_c5_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_c5_pytest.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/r02_c5_bench.json 2> gpurun_out/r02_c5_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c5_bench.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['block_roofline']['frac'], d['block_roofline']['ms'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'])
for c in d['roofline']['by_class']: print(c)
PY
