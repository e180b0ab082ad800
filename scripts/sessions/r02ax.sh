#!/bin/bash
# round 2, GPU session AX: set_points with copy and upload overlapped (tm_mesh_set_positions_rows), e2e headline through it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_set_points.py tests/test_gpu_energy.py -m gpu -q -x 2>&1 | tail -3
for t in 0 2 1; do TM_SET_POINTS_THREADS=$t timeout 300 python scripts/set_points_timing.py 8 2>&1 | tail -1; done
timeout 1500 python bench.py --no-partition --replicas 0 --no-hmc-step > gpurun_out/r02ax_bench.json 2> gpurun_out/r02ax_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02ax_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')})
e = d['e2e']; print(e['value'], e['ms_per_step'], e['phase_ms']); print(e['reference_idiom']); print(e['host'])
PY
tail -3 gpurun_out/r02ax_bench.err
