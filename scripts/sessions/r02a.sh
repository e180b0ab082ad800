#!/bin/bash
# round 2, GPU session A: parity of the closed-fan sweeps, kernel timings of the register/occupancy variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r02a_gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02a_tests.log
tail -5 gpurun_out/r02a_tests.log
./scripts/ubench/seed_accuracy > gpurun_out/r02a_seed_accuracy.txt 2>&1; cat gpurun_out/r02a_seed_accuracy.txt
: > gpurun_out/r02a_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02a_kernels.jsonl 2>gpurun_out/r02a_kb.err
for v in g10 g12 g14 p10 p14 p16 nt128; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02a_kernels.jsonl 2>>gpurun_out/r02a_kb.err
done
cat gpurun_out/r02a_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') or k.endswith('_ms')}, d['E'], d['gnorm'])
"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 1500 gpurun_out/r02a_bench.json
