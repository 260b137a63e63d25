#!/bin/bash
# space attention backward, split-phase schedule: correctness vs the mma.sync kernels, timing, cycle stamps, attention / block tests
timeout 120 python tools/gpu_attn_tc_probe.py bwd 64 12 16 196 2>&1 | tail -7
timeout 120 python tools/gpu_cls_fusion_probe.py 64 12 16 196 0 2>&1 | tail -6
timeout 120 python tools/gpu_attn_bwd_stamps.py 2>&1 | sed -n 1,8p
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider -k "attn or attention or block" 2>&1 | tail -2
