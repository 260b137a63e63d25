#!/bin/bash
timeout 600 python tools/gpu_narrator_profile.py 10 24 2>&1 | grep -v "^=>\|Warning" | tail -50
