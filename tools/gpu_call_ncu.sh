mkdir -p gpurun_out
timeout 170 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"space_attn_bwd_tc|gemm2_bf16" -c 12 -o gpurun_out/ncu_r01_final_bwd -f python tools/profile_step.py --batch 16 --range bwd > gpurun_out/ncu_final_bwd.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/ncu_r01_final_bwd.ncu-rep
