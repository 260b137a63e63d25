mkdir -p gpurun_out
export LAVILA_B200_DECODE_GRAPH=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_narrator.csv python tools/bench_narrator.py --encoder large --batch 32 --returns 1 --max-len 6 > gpurun_out/ncu_narr.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_narr.log | cut -c1-300
