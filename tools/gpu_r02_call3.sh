#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_fulldepth_diag.py 12 3 1 > gpurun_out/r02_c3_diag12.log 2>&1; grep -v "blocks\.[2-9]\.\|blocks\.10\|resblocks\.[1-9]" gpurun_out/r02_c3_diag12.log | tail -60
timeout 300 python tools/gpu_fulldepth_diag.py 2 3 1 > gpurun_out/r02_c3_diag2.log 2>&1; tail -50 gpurun_out/r02_c3_diag2.log | grep -v "resblocks\.[1-9]"
