#!/bin/bash
# round 2, GPU session F: direct register-gather block walks
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02f_tests.log
tail -5 gpurun_out/r02f_tests.log
: > gpurun_out/r02f_kernels.jsonl
TRIMEM_B200_NO_TILES=1 timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02f_kernels.jsonl 2>gpurun_out/r02f_kb.err
for v in b4 b6 g2 g8b4 n64b10 n64b12 n256b2; do
  TRIMEM_B200_NO_TILES=1 TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02f_kernels.jsonl 2>>gpurun_out/r02f_kb.err
done
cat gpurun_out/r02f_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') or k.endswith('_ms')}, d['E'], d['gnorm'])
"
TRIMEM_B200_NO_TILES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_props|k_grad' -c 4 -o gpurun_out/r02f_sweeps python scripts/prof_kernels.py 8 stress 1 2 > gpurun_out/r02f_ncu.log 2>&1
ls -la gpurun_out/r02f*.ncu-rep
python scripts/prof_pflip.py > gpurun_out/r02f_pflip.log 2>&1; tail -3 gpurun_out/r02f_pflip.log
