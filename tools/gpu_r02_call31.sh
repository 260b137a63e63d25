#!/bin/bash
timeout 200 python tools/gpu_skinny_tc_probe.py 2>&1 | tail -9
LAVILA_B200_SKINNY_TC=0 timeout 200 python tools/gpu_skinny_tc_probe.py 2>&1 | tail -9
