mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 tools/gpu_p2p_loss_check.py > gpurun_out/p2p_check4.log 2>&1; echo "p2p check (4 GPUs) rc=$?"; grep -v "^\s*$" gpurun_out/p2p_check4.log | grep -v Warning | tail -6 | cut -c1-600
