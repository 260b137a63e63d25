# GPU call: key-tiled attention + SSL loss parity, then TSF-L narrator / training timings
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "tiled_group" -p no:cacheprovider > gpurun_out/t_tiled.log 2>&1; echo "tiled rc=$?"; tail -30 gpurun_out/t_tiled.log
timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -k "var_attention or ssl or tsf_l14" -p no:cacheprovider > gpurun_out/t_rest.log 2>&1; echo "rest rc=$?"; tail -30 gpurun_out/t_rest.log
timeout 300 python tools/bench_narrator.py --encoder large --batch 32 --returns 1 > gpurun_out/narrator_large.json 2> gpurun_out/narrator_large.err; echo "narrator rc=$?"; tail -3 gpurun_out/narrator_large.err; cat gpurun_out/narrator_large.json
timeout 300 python bench.py --model large336 --frames 32 --batch 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_large336.json 2> gpurun_out/bench_large336.err; echo "large336 rc=$?"; tail -3 gpurun_out/bench_large336.err; cat gpurun_out/bench_large336.json
