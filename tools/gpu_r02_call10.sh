#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider ) > gpurun_out/r02_c10_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_c10_pytest.log | cut -c1-300
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02_c10_bench.json 2> gpurun_out/r02_c10_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r02_c10_bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['block_roofline']['frac'], d['roofline']['frac'], d.get('vs_eager'), d.get('cpu_baseline'))
        print(d.get('narrator'))
PY
tail -4 gpurun_out/r02_c10_bench.err
