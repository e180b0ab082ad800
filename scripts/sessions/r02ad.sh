#!/bin/bash
# round 2, GPU session AD: cell-list pair sweep, a warp per vertex: timing at 655,362 and 40,962 vertices against the thread version; tests
mkdir -p gpurun_out
: > gpurun_out/r02ad_kernels.jsonl
for sub in 8 6; do
  timeout 300 python scripts/kernel_bench.py $sub stress cell-list >> gpurun_out/r02ad_kernels.jsonl 2>gpurun_out/r02ad_kb.err
  TRIMEM_B200_LIB=trimem_b200/lib/variants/cellthread.so timeout 300 python scripts/kernel_bench.py $sub stress cell-list >> gpurun_out/r02ad_kernels.jsonl 2>>gpurun_out/r02ad_kb.err
done
cat gpurun_out/r02ad_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['V'], d['lib'][-14:], {k: d[k] for k in d if k.endswith('_us') and 'warm' not in k}, d['E'], d['gnorm'])
"
tail -3 gpurun_out/r02ad_kb.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02ad_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02ad_tests.log
tail -4 gpurun_out/r02ad_tests.log
