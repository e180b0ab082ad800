#!/bin/bash
# round 2, GPU session B: tile-staged ring sweeps: parity, variants, one ncu --set full capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02b_tests.log
tail -5 gpurun_out/r02b_tests.log
: > gpurun_out/r02b_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02b_kernels.jsonl 2>gpurun_out/r02b_kb.err
for v in n64 n128b6 n128b4 n256b3 n256b2; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02b_kernels.jsonl 2>>gpurun_out/r02b_kb.err
done
cat gpurun_out/r02b_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') or k.endswith('_ms')}, d['E'], d['gnorm'])
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_props|k_grad|k_pairs' -c 9 -o gpurun_out/r02b_sweeps python scripts/prof_kernels.py > gpurun_out/r02b_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
