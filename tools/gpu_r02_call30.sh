#!/bin/bash
# decode-sized GEMMs with more than 64 rows on tcgen05 (swap-AB): parity, narrator tests, config 4 throughput A/B
timeout 600 python -m pytest tests/test_gpu_narrator.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/bench_narrator.py --encoder large --batch 32 --returns 10,4 2>&1 | tail -2
LAVILA_B200_SKINNY_TC=0 timeout 300 python tools/bench_narrator.py --encoder large --batch 32 --returns 10 2>&1 | tail -1
