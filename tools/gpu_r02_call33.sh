#!/bin/bash
# 2 GPUs: the fused gather + loss test, then a short 2-GPU bench (loss_check + DDP exposure)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_r02_final2_2gpu.json 2> gpurun_out/bench_r02_final2_2gpu.err
echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r02_final2_2gpu.err; head -c 600 gpurun_out/bench_r02_final2_2gpu.json
