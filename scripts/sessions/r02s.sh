#!/bin/bash
# round 2, GPU session S: pair sweep occupancy variants after the integer guards; full GPU tests; ncu of the pair sweep
mkdir -p gpurun_out
: > gpurun_out/r02s_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02s_kernels.jsonl 2>gpurun_out/r02s_kb.err
for v in qpb2b8 qpb1b8 qpb3b7; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02s_kernels.jsonl 2>>gpurun_out/r02s_kb.err
done
cat gpurun_out/r02s_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') and 'warm' not in k}, d['E'], d['gnorm'])
"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02s_tests.log
tail -4 gpurun_out/r02s_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pairs' -c 2 --launch-skip 2 -o gpurun_out/r02s_pairs python scripts/prof_kernels.py 8 stress 3 > gpurun_out/r02s_ncu.log 2>&1
ls -la gpurun_out/r02s_pairs.ncu-rep
