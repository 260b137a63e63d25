#!/bin/bash
mkdir -p gpurun_out
echo "== rows ON"; timeout 120 python tools/gpu_gemm2_epi_probe.py all 20
echo "== rows OFF"; LAVILA_B200_GEMM_ROWS_EPI=0 timeout 120 python tools/gpu_gemm2_epi_probe.py all 20
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm2_bf16_kernel" -s 2 -c 1 -o gpurun_out/ncu_r02_gemm2_fc1_rows -f python tools/gpu_gemm2_epi_probe.py fc1 1 > gpurun_out/r02_c7_ncu_fc1.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm2_bf16_kernel" -s 2 -c 1 -o gpurun_out/ncu_r02_gemm2_dgelu_rows -f python tools/gpu_gemm2_epi_probe.py dgelu 1 > gpurun_out/r02_c7_ncu_dgelu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
