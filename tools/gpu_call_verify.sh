mkdir -p gpurun_out
( time timeout 420 python -m pytest tests -m gpu -x -q ) > gpurun_out/gpu_tests_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_full.log
tail -5 gpurun_out/gpu_tests_full.log
timeout 300 python bench.py > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; echo "bench rc=$?"
cat gpurun_out/bench_head.json
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_head.csv python tools/profile_step.py --batch 64 --range step > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
