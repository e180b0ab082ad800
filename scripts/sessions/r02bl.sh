#!/bin/bash
# round 2, GPU session BL: falling priorities along the side streams: 655,362 and 10,485,762 vertices
mkdir -p gpurun_out
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['V'], {k: d[k] for k in d if k.endswith('_us') and 'warm' not in k})"
timeout 900 python bench.py --steps 10 --replicas 0 --no-hmc-step --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['partitioned']; print(d['value'], d['leapfrog']['ms_per_step'], '1 rank:', {k: p.get(k) for k in ('ranks','ms_per_eg','leapfrog_ms_per_step')})"
