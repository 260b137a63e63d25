mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/gpu_p2p_loss_check.py > gpurun_out/p2p_check.log 2>&1; echo "p2p check rc=$?"; tail -15 gpurun_out/p2p_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; tail -5 gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json
