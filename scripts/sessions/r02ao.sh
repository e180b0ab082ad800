#!/bin/bash
# round 2, GPU session AO: pair sweep cut into K tile ranges on K side streams, each range's ring gradient behind it (timing experiment)
mkdir -p gpurun_out
: > gpurun_out/r02ao_kernels.jsonl
for v in split0 split2 split4 split8; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02ao_kernels.jsonl 2>>gpurun_out/r02ao_kb.err
done
cat gpurun_out/r02ao_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-10:], {k: d[k] for k in d if k.endswith('_us') and 'warm' not in k}, d['E'], d['gnorm'])
"
tail -3 gpurun_out/r02ao_kb.err
