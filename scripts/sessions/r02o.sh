#!/bin/bash
# round 2, GPU session O: pair sweep split into head and tail, ring gradient of the head beside the tail: share sweep, GPU tests
mkdir -p gpurun_out
: > gpurun_out/r02o_kernels.jsonl
for sp in 0 50 60 70 80 90; do
  TRIMEM_B200_PAIR_SPLIT=$sp timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02o_kernels.jsonl 2>>gpurun_out/r02o_kb.err
done
cat gpurun_out/r02o_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in d if k.endswith('_us') and 'warm' not in k}, d['E'], d['gnorm'], d['E_core'], d['gnorm_core'])
"
tail -3 gpurun_out/r02o_kb.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02o_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02o_tests.log
tail -4 gpurun_out/r02o_tests.log
