#!/bin/bash
# 2 GPUs: DDP all-reduce exposure with the work-stealing tile scheduler of the 2-CTA GEMM on / off
for dyn in 1 0; do
LAVILA_B200_GEMM_DYN_SCHED=$dyn timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$dyn bench.py --gpus 2 --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('DYN=$dyn', d['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d['ddp']), d['loss_check']['ok'], d['clocks']['sm_mhz'])"
done
