#!/bin/bash
# round 2, GPU session L: tau-early / occupancy variants of the ring gradient, full GPU tests on the current build, bench line
mkdir -p gpurun_out
: > gpurun_out/r02l_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02l_kernels.jsonl 2>gpurun_out/r02l_kb.err
for v in te0b5 te0b6 te1b4 te1b4g8 te1b5g2 te1b5g8; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02l_kernels.jsonl 2>>gpurun_out/r02l_kb.err
done
cat gpurun_out/r02l_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') or k.endswith('_ms')}, d['E'], d['gnorm'])
"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02l_tests.log
tail -4 gpurun_out/r02l_tests.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02l_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}); print(d['roofline']['kernel_ms'], d['roofline']['eg_frac_of_peak'], d['roofline']['leapfrog_frac_of_peak']); print(d['leapfrog']['ms_per_step'], d['verlet_rebuild_ms'], d['e2e']['value'])
PY
