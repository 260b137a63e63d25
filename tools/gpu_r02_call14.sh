#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gpu_cls_fusion_probe.py 2 2 4 196 2>&1 | tail -12
timeout 120 python tools/gpu_cls_fusion_probe.py 64 12 16 196 2>&1 | tail -12
( timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider ) > gpurun_out/r02_c14_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_c14_pytest.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline --no-narrator > gpurun_out/r02_c14_bench.json 2> gpurun_out/r02_c14_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c14_bench.json').read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['block_roofline']['frac'], d['block_roofline']['ms'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['clocks'])
PY
