#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_step_b64.csv python tools/profile_step.py --batch 64 --range step > gpurun_out/r02_c22_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarise_launches.py gpurun_out/launches_r02_step_b64.csv 32 | cut -c1-140
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02_c22_bench.json 2> gpurun_out/r02_c22_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r02_c22_bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['block_roofline']['frac'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d.get('vs_eager'), d['eager_baseline']['value'], d['clocks'])
        print(d.get('narrator'))
PY
