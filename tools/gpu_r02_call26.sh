#!/bin/bash
# input pipeline: parity tests, the bench leg, one ncu capture of the kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_input_pipeline.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python tools/gpu_input_pipeline_probe.py 2>&1 | tail -3 | tee gpurun_out/input_pipeline_r02.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:clip_transform -c 2 -o gpurun_out/ncu_r02_clip_transform \
  python tools/gpu_input_pipeline_probe.py > gpurun_out/ncu_clip_transform.log 2>&1
ls -la gpurun_out | tail -5
