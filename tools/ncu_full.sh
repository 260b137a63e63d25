#!/bin/bash
# ncu --set full captures of the heavy kernels (B=16 keeps replay time low; same kernel structure as B=64).
mkdir -p gpurun_out
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"group_attn_bwd|cls_attn_bwd|ln_bwd" -c 9 -o gpurun_out/ncu_r01_bwd -f \
    python tools/profile_step.py --batch 16 --range bwd > gpurun_out/ncu_bwd.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"group_attn_fwd|gemm_bf16_kernel|cls_attn_fwd|ln_fwd" -c 16 -o gpurun_out/ncu_r01_fwd -f \
    python tools/profile_step.py --batch 16 --range fwd > gpurun_out/ncu_fwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
