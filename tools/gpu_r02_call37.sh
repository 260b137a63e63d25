#!/bin/bash
# smoke() of the final state + one ncu --set full capture of the time attention (fused CLS query) forward / backward at batch 64
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:time_attn -c 4 -o gpurun_out/ncu_r02_time_attn_b64 \
  python tools/gpu_cls_fusion_probe.py 64 12 16 196 1 > gpurun_out/ncu_time_attn.log 2>&1
tail -3 gpurun_out/ncu_time_attn.log; ls -la gpurun_out/*.ncu-rep | tail -2
