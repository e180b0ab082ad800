#!/bin/bash
# round 2, GPU session M: ring gradient with cp.async-staged ring records: variants, then the GPU tests on the default
mkdir -p gpurun_out
: > gpurun_out/r02m_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02m_kernels.jsonl 2>gpurun_out/r02m_kb.err
for v in gs0 gs1b6 gs1b6g8 gs1b5g8 gs1b6g2 gs1b8; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02m_kernels.jsonl 2>>gpurun_out/r02m_kb.err
done
cat gpurun_out/r02m_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') or k.endswith('_ms')}, d['E'], d['gnorm'])
"
tail -3 gpurun_out/r02m_kb.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02m_tests.log
tail -4 gpurun_out/r02m_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_grad' -c 3 --launch-skip 4 -o gpurun_out/r02m_grad python scripts/prof_kernels.py 8 stress 2 1 > gpurun_out/r02m_ncu.log 2>&1
ls -la gpurun_out/r02m_grad.ncu-rep
