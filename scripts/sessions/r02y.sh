#!/bin/bash
# round 2, GPU session Y: final build: GPU tests, the default bench line, the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02y_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02y_tests.log
tail -3 gpurun_out/r02y_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r02y_bench.json 2> gpurun_out/r02y_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02y_bench_ref.json 2> gpurun_out/r02y_bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r02y_bench_ref.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02y_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}); print(d['roofline']['kernel_ms'], d['roofline']['eg_frac_of_peak'], d['roofline']['leapfrog_frac_of_peak'], d['roofline']['frac'])
print(d['leapfrog']['ms_per_step'], d['verlet_rebuild_ms'], d['e2e']['value'], d['e2e']['hmc_move']['ms_per_move'])
print(d['replicas']); print(d['hmc_step_config2']); print(d['partitioned'])
PY
