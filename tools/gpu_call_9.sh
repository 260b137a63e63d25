mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_narrator.py -q -p no:cacheprovider > gpurun_out/t9.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/t9.log | cut -c1-300
timeout 400 python tools/bench_narrator.py --encoder large --batch 32 --returns 1,10 > gpurun_out/narrator_large_skinny.json 2> gpurun_out/narrator_large_skinny.err; echo "narrator rc=$?"; tail -3 gpurun_out/narrator_large_skinny.err | cut -c1-300; cat gpurun_out/narrator_large_skinny.json
