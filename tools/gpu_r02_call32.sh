#!/bin/bash
# full GPU test suite + the default bench line (round-2 state: gradient pool, input pipeline leg)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
timeout 1200 python bench.py > gpurun_out/bench_r02_final2_b64.json 2> gpurun_out/bench_r02_final2_b64.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r02_final2_b64.err; head -c 1500 gpurun_out/bench_r02_final2_b64.json
