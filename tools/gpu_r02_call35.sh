#!/bin/bash
# dynamic tile scheduling in the 2-CTA GEMM: GEMM parity (unit + fuzz + model), then the step with the scheduler on / off
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fuzz_vs_doubles.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 400 python bench.py --steps 6 --warmup 3 --no-eager-baseline --no-narrator --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('DYN on ', d['value'], d['ms_per_step'], d['roofline']['frac'], d['block_roofline']['frac'], d['clocks']['sm_mhz'])"
LAVILA_B200_GEMM_DYN_SCHED=0 timeout 400 python bench.py --steps 6 --warmup 3 --no-eager-baseline --no-narrator --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('DYN off', d['value'], d['ms_per_step'], d['roofline']['frac'], d['block_roofline']['frac'], d['clocks']['sm_mhz'])"
