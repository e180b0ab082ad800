#!/bin/bash
# round 2, GPU session AW: phases of the end-to-end step, copyto vs set_points, host threads
for t in 4 1 8 16; do TM_SET_POINTS_THREADS=$t timeout 300 python scripts/set_points_timing.py 8 2>&1 | tail -2; done
nproc; lscpu | grep -i "numa\|model name\|socket" | head -8
