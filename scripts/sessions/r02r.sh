#!/bin/bash
# round 2, GPU session R: pair sweep written for instruction count (integer guards, ping-pong record buffers): variants
mkdir -p gpurun_out
: > gpurun_out/r02r_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02r_kernels.jsonl 2>gpurun_out/r02r_kb.err
for v in qil qpb2 qpb4b5 qpb2b7; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02r_kernels.jsonl 2>>gpurun_out/r02r_kb.err
done
cat gpurun_out/r02r_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') and 'warm' not in k}, d['E'], d['gnorm'])
"
tail -3 gpurun_out/r02r_kb.err
timeout 600 python -m pytest tests/test_gpu_energy.py tests/test_gpu_neighbours.py -m gpu -q -x 2>&1 | tail -5
