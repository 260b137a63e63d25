#!/bin/bash
# round 2, call 1: parity suite (incl. the new real-size tests), the never-run fuzz sweeps, bench with the eager baseline,
# ncu --set full of the tcgen05 space attention at B=64
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r02_c1_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_c1_pytest.log | cut -c1-300
( time timeout 600 python -m pytest tests/test_gpu_fuzz_vs_doubles.py -m gpu_fuzz -q -p no:cacheprovider ) > gpurun_out/r02_c1_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -25 gpurun_out/r02_c1_fuzz.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_c1_bench.json 2> gpurun_out/r02_c1_bench.err; echo "bench rc=$?"; cat gpurun_out/r02_c1_bench.json | cut -c1-4000; tail -5 gpurun_out/r02_c1_bench.err
timeout 200 python bench.py --impl eager --amp fp16 --steps 4 --warmup 2 > gpurun_out/r02_c1_eager_fp16.json 2> gpurun_out/r02_c1_eager_fp16.err; echo "eager fp16 rc=$?"; cat gpurun_out/r02_c1_eager_fp16.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"space_attn_bwd_tc" -c 2 -o gpurun_out/ncu_r02_attn_tc_bwd_b64 -f \
    python tools/gpu_attn_tc_probe.py bwd 64 12 16 196 > gpurun_out/r02_c1_ncu_attn_bwd.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_c1_ncu_attn_bwd.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"space_attn_fwd_tc" -c 2 -o gpurun_out/ncu_r02_attn_tc_fwd_b64 -f \
    python tools/gpu_attn_tc_probe.py fwd 64 12 16 196 > gpurun_out/r02_c1_ncu_attn_fwd.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_c1_ncu_attn_fwd.log
ls -la gpurun_out/*.ncu-rep
