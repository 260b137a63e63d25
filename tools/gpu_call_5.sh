mkdir -p gpurun_out
for v in 0 1 0 1; do
  export LAVILA_B200_WGRAD_STREAM=$v
  timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline 2>gpurun_out/ab_$v.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('WGRAD_STREAM=$v', d['value'], d['ms_per_step'], d['clocks'])"
done
tail -3 gpurun_out/ab_1.err
