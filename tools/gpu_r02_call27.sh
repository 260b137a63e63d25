#!/bin/bash
# input pipeline (4 px / thread kernel) + space attention backward with double-buffered TMEM loads in the elementwise warps
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_input_pipeline.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
timeout 300 python tools/gpu_input_pipeline_probe.py 2>&1 | tail -1 | tee gpurun_out/input_pipeline_r02.json
timeout 120 python tools/gpu_attn_tc_probe.py bwd 64 12 16 196 2>&1 | tail -7
timeout 120 python tools/gpu_cls_fusion_probe.py 64 12 16 196 0 2>&1 | tail -6
timeout 120 python tools/gpu_attn_bwd_stamps.py 2>&1 | sed -n 1,8p
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
