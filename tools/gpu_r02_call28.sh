#!/bin/bash
# space attention backward, pipelined half-step schedule (PIPE): correctness vs the mma.sync kernels, timing vs the serial schedule
timeout 120 python tools/gpu_attn_tc_probe.py bwd 64 12 16 196 2>&1 | tail -7
timeout 120 python tools/gpu_cls_fusion_probe.py 64 12 16 196 0 2>&1 | tail -6
timeout 120 python tools/gpu_attn_bwd_stamps.py 2>&1 | sed -n 1,8p
LAVILA_B200_ATTN_BWD_PIPE=0 timeout 120 python tools/gpu_attn_tc_probe.py bwd 64 12 16 196 2>&1 | grep "bwd tc=True"
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "attn or attention or block" 2>&1 | tail -2
