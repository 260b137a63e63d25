#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --model large336 --frames 32 --batch 4 --steps 6 --warmup 3 --no-e2e > gpurun_out/r02_large336_b4.json 2> gpurun_out/r02_large336_b4.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_large336_b4.json').read()); print(d['metric'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['max_mem_gb'])
PY
LAVILA_B200_FLASH_TC=0 timeout 600 python bench.py --model large336 --frames 32 --batch 4 --steps 6 --warmup 3 --no-e2e --no-roofline > gpurun_out/r02_large336_b4_notc.json 2> /dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_large336_b4_notc.json').read()); print("mma.sync fwd:", d['value'], d['ms_per_step'])
PY
