#!/bin/bash
# round 2, GPU session C: paired-slot walks, tile pair lists: parity, variants, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_flip_statistics.py > gpurun_out/r02c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02c_tests.log
tail -5 gpurun_out/r02c_tests.log
: > gpurun_out/r02c_kernels.jsonl
TRIMEM_B200_DEBUG=1 timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02c_kernels.jsonl 2>gpurun_out/r02c_kb.err
grep "tile lists" gpurun_out/r02c_kb.err | head -2
TRIMEM_B200_NO_TILES=1 timeout 300 python scripts/kernel_bench.py 8 stress verlet-list | sed 's/"lib": "default"/"lib": "notiles"/' >> gpurun_out/r02c_kernels.jsonl 2>>gpurun_out/r02c_kb.err
for v in b4 b6 n256b2 n256b3 n64b8 nocubic; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02c_kernels.jsonl 2>>gpurun_out/r02c_kb.err
done
cat gpurun_out/r02c_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') or k.endswith('_ms')}, d['E'], d['gnorm'])
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_props|k_grad|k_pairs' -c 8 -o gpurun_out/r02c_sweeps python scripts/prof_kernels.py > gpurun_out/r02c_ncu.log 2>&1
ls -la gpurun_out/r02c*.ncu-rep
