mkdir -p gpurun_out
( time timeout 500 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/gpu_tests_final.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/gpu_tests_final.log | cut -c1-200
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cat gpurun_out/bench_final.json | cut -c1-2500
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python tools/profile_step.py --batch 64 --range step > gpurun_out/ncu_launch_final.log 2>&1; echo "ncu rc=$?"
