#!/bin/bash
# 4 GPUs: short bench (loss_check of the fused gather at 4 ranks / 256 rows, DDP exposure)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 4 --warmup 3 2>/dev/null | tee gpurun_out/bench_r02_final2_4gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], json.dumps(d['loss_check']), json.dumps(d['ddp']), d['clocks']['sm_mhz'])"
