#!/bin/bash
# launch list of one full step (round-2 final state) under ncu, one metric, cudaProfilerStart/Stop around the step
mkdir -p gpurun_out
timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_final2_step_b64.csv python tools/profile_step.py --batch 64 --range step > gpurun_out/r02_c34_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r02_c34_ncu.log; wc -l gpurun_out/launches_r02_final2_step_b64.csv
