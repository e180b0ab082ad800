#!/bin/bash
# round 2, GPU session D: block walks (whole ring in one basic block), pipelined tile pair sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_flip_statistics.py > gpurun_out/r02d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02d_tests.log
tail -5 gpurun_out/r02d_tests.log
: > gpurun_out/r02d_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02d_kernels.jsonl 2>gpurun_out/r02d_kb.err
TRIMEM_B200_NO_TILES=1 timeout 300 python scripts/kernel_bench.py 8 stress verlet-list | sed 's/"lib": "default"/"lib": "notiles"/' >> gpurun_out/r02d_kernels.jsonl 2>>gpurun_out/r02d_kb.err
for v in g2 g6b4 g3b4 g3b6 g2b6 pt3 n256; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02d_kernels.jsonl 2>>gpurun_out/r02d_kb.err
done
cat gpurun_out/r02d_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') or k.endswith('_ms')}, d['E'], d['gnorm'])
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_props|k_grad|k_pairs' -c 8 -o gpurun_out/r02d_sweeps python scripts/prof_kernels.py 8 stress 1 2 3 > gpurun_out/r02d_ncu.log 2>&1
ls -la gpurun_out/r02d*.ncu-rep
