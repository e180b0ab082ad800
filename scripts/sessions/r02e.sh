#!/bin/bash
# round 2, GPU session E: persistent pipelined ring sweeps
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_flip_statistics.py > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02e_tests.log
tail -5 gpurun_out/r02e_tests.log
: > gpurun_out/r02e_kernels.jsonl
TRIMEM_B200_DEBUG=1 TRIMEM_B200_NO_TILES=1 timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02e_kernels.jsonl 2>gpurun_out/r02e_kb.err
grep "ring sweeps" gpurun_out/r02e_kb.err | head -1
for v in ks1 ks3 ks2b4 ks2b6 ks1b6 n64 n256; do
  TRIMEM_B200_NO_TILES=1 TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02e_kernels.jsonl 2>>gpurun_out/r02e_kb.err
done
cat gpurun_out/r02e_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') or k.endswith('_ms')}, d['E'], d['gnorm'])
"
TRIMEM_B200_NO_TILES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_props|k_grad' -c 4 -o gpurun_out/r02e_sweeps python scripts/prof_kernels.py 8 stress 1 2 > gpurun_out/r02e_ncu.log 2>&1
ls -la gpurun_out/r02e*.ncu-rep
timeout 300 python -m pytest tests/test_gpu_flip_statistics.py -m gpu -q -x > gpurun_out/r02e_flipstats.log 2>&1; tail -15 gpurun_out/r02e_flipstats.log
