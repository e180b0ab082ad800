#!/bin/bash
# round 2, GPU session H: pybind partition test, pair-sweep residency cap, full bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02h_tests.log
tail -5 gpurun_out/r02h_tests.log
: > gpurun_out/r02h_kernels.jsonl
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02h_kernels.jsonl 2>gpurun_out/r02h_kb.err
for v in pad47 pad70 pad100; do
  TRIMEM_B200_LIB=trimem_b200/lib/variants/$v.so timeout 300 python scripts/kernel_bench.py 8 stress verlet-list >> gpurun_out/r02h_kernels.jsonl 2>>gpurun_out/r02h_kb.err
done
cat gpurun_out/r02h_kernels.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'][-12:], {k: d[k] for k in d if k.endswith('_us') or k.endswith('_ms')}, d['E'], d['gnorm'])
"
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02h_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02h_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}); print(d['e2e']); print(d['replicas']); print(d['roofline']); print(d['leapfrog']); print(d['partitioned'])
PY
