#!/bin/bash
mkdir -p gpurun_out
for k in 0 4 8; do
LAVILA_B200_RESERVE_SMS=$k timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e --no-roofline > gpurun_out/r02_2gpu_k$k.json 2> gpurun_out/r02_2gpu_k$k.err; echo "k=$k rc=$?"; python - <<PY
import json
for l in open('gpurun_out/r02_2gpu_k$k.json'):
    if l.startswith('{'):
        d=json.loads(l); print("reserve=$k", d['value'], d['ms_per_step'], d.get('ddp'), d.get('loss_check',{}).get('ok'))
PY
done
