#!/bin/bash
# round 2, GPU session X: ring gradient: sums requested with the row, tether slots from the register row; occupancy / fence variants; ncu
mkdir -p gpurun_out
: > gpurun_out/r02x_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02x_kernels.jsonl 2>gpurun_out/r02x_kb.err
for v in gb4g8 gb5 gb3g8 gb4g2; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02x_kernels.jsonl 2>>gpurun_out/r02x_kb.err
done
cat gpurun_out/r02x_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') and 'warm' not in k}, d['E'], d['gnorm'])
"
tail -3 gpurun_out/r02x_kb.err

timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_grad' -c 2 --launch-skip 1 -o gpurun_out/r02x_grad python scripts/prof_kernels.py 8 stress 2 >> gpurun_out/r02x_ncu.log 2>&1
ls -la gpurun_out/r02x_grad*.ncu-rep
