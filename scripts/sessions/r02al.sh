#!/bin/bash
# round 2, GPU session AL: the pair sweep adds up its own repulsion sum (no join in front of the property reduction): tests + timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02al_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02al_tests.log
tail -5 gpurun_out/r02al_tests.log
: > gpurun_out/r02al_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02al_kernels.jsonl 2>gpurun_out/r02al_kb.err
for v in prio_a; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02al_kernels.jsonl 2>>gpurun_out/r02al_kb.err
done
cat gpurun_out/r02al_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') and 'warm' not in k}, d['E'], d['gnorm'])
"
tail -3 gpurun_out/r02al_kb.err
