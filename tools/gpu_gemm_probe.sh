#!/bin/bash
# Runs every GEMM probe variant in its own process; logs to gpurun_out/gemm_probe.log
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for v in "nt 1000 768 768" "nt 128 256 64" "nt 4096 2304 768" "nn 1000 768 768" "nn 128 256 64" "tn 768 768 1024" "tn 128 256 64" "tk 768 768 1024" "splitk 768 768 6400" "epi 1000 768 512" "perf"; do
  echo "=== $v"
  timeout 120 python tools/gpu_gemm_probe.py $v 2>&1 | tail -15
done
} 2>&1 | tee gpurun_out/gemm_probe.log
