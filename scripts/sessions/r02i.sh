#!/bin/bash
# round 2, GPU session I: occupancy variants of the final kernels
mkdir -p gpurun_out
: > gpurun_out/r02i_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02i_kernels.jsonl 2>gpurun_out/r02i_kb.err
for v in p7 p8 g6 g7 g8 q7 q8b2 q6b2; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02i_kernels.jsonl 2>>gpurun_out/r02i_kb.err
done
cat gpurun_out/r02i_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') or k.endswith('_ms')}, d['E'], d['gnorm'])
"
