#!/bin/bash
# round 2, GPU session AQ: split pair sweep, repulsion column reduced behind the join (no tail in the sweep): tests + timings by K
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pair_split.py -m gpu -q -x > gpurun_out/r02aq_tests.log 2>&1; echo "split tests rc=$?" >> gpurun_out/r02aq_tests.log
tail -25 gpurun_out/r02aq_tests.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02aq_tests_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02aq_tests_all.log
tail -8 gpurun_out/r02aq_tests_all.log
: > gpurun_out/r02aq_kernels.jsonl
for k in 2 0 3 4; do
TM_PAIR_SPLIT=$k timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02aq_kernels.jsonl 2>>gpurun_out/r02aq_kb.err
done
TM_PAIR_SPLIT=2 timeout 300 python scripts/kernel_bench.py 7 stress verlet-list >> gpurun_out/r02aq_kernels.jsonl 2>>gpurun_out/r02aq_kb.err
TM_PAIR_SPLIT=0 timeout 300 python scripts/kernel_bench.py 7 stress verlet-list >> gpurun_out/r02aq_kernels.jsonl 2>>gpurun_out/r02aq_kb.err
cat gpurun_out/r02aq_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['V'], {k: d[k] for k in d if k.endswith('_us') and 'warm' not in k}, d['E'], d['gnorm'])
"
tail -3 gpurun_out/r02aq_kb.err
