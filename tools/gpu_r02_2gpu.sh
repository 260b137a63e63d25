#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_2gpu_bench.json 2> gpurun_out/r02_2gpu_bench.err; echo "bench rc=$?"; cat gpurun_out/r02_2gpu_bench.json | cut -c1-3000; tail -15 gpurun_out/r02_2gpu_bench.err | cut -c1-400
