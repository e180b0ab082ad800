#!/bin/bash
# round 2, GPU session N: cold vs warm L2 per kernel; L2 fill granularity of the record gathers; ncu of the staged ring gradient
mkdir -p gpurun_out
: > gpurun_out/r02n_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02n_kernels.jsonl 2>gpurun_out/r02n_kb.err
for v in gs0 pf64s0 pf128s0 pf256s0 pf128s1; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02n_kernels.jsonl 2>>gpurun_out/r02n_kb.err
done
cat gpurun_out/r02n_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') or k.endswith('_ms')}, d['E'], d['gnorm'])
"
tail -3 gpurun_out/r02n_kb.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_grad' -c 2 --launch-skip 1 -o gpurun_out/r02n_grad python scripts/prof_kernels.py 8 stress 2 > gpurun_out/r02n_ncu.log 2>&1
ls -la gpurun_out/r02n_grad.ncu-rep
