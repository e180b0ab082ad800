#!/bin/bash
# round 2, GPU session AN: tile protocol with one release-atomic per block instead of fences: tests + timings with / without
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair_dec.py tests/test_gpu_trajectory.py tests/test_gpu_energy.py -m gpu -q -x > gpurun_out/r02an_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02an_tests.log
tail -5 gpurun_out/r02an_tests.log
: > gpurun_out/r02an_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02an_kernels.jsonl 2>gpurun_out/r02an_kb.err
TM_STREAM_PRIO=0 timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02an_kernels.jsonl 2>>gpurun_out/r02an_kb.err
TM_PAIR_DEC=0 timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02an_kernels.jsonl 2>>gpurun_out/r02an_kb.err
cat gpurun_out/r02an_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in d if k.endswith('_us') and 'warm' not in k}, d['E'], d['gnorm'])
"
tail -3 gpurun_out/r02an_kb.err
