#!/bin/bash
# round 2, GPU session AR: bench line of the split pair sweep build, reference arm
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r02ar_bench.json 2> gpurun_out/r02ar_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02ar_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}); print(d['roofline']['kernel_ms'], d['roofline']['eg_frac_of_peak'], d['roofline']['leapfrog_frac_of_peak'], d['roofline']['frac'])
print(d['leapfrog']['ms_per_step'], d['verlet_rebuild_ms'], d['e2e']['value'], d['e2e']['phase_ms'], d['e2e']['hmc_move']['ms_per_move'])
print(d['replicas']); print(d['hmc_step_config2']); print(d['partitioned'])
PY
tail -5 gpurun_out/r02ar_bench.err
