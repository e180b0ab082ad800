#!/bin/bash
# round 2, GPU session AK: pair sweep decoupled from the ring gradient (timing only: joins deferred), stream priorities
mkdir -p gpurun_out
: > gpurun_out/r02ak_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02ak_kernels.jsonl 2>gpurun_out/r02ak_kb.err
for v in nojoin nojoin_pa nojoin_pb prio_a prio_b; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02ak_kernels.jsonl 2>>gpurun_out/r02ak_kb.err
done
cat gpurun_out/r02ak_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') and 'warm' not in k}, d['E'], d['gnorm'])
"
tail -3 gpurun_out/r02ak_kb.err
