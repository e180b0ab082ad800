#!/bin/bash
# round 2, GPU session AV: TriMesh.set_points (one upload per position change), host copy / PCIe rates in the bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_set_points.py tests/test_gpu_energy.py tests/test_gpu_trajectory.py -m gpu -q -x 2>&1 | tail -4
timeout 1500 python bench.py --no-partition --replicas 0 --no-hmc-step > gpurun_out/r02av_bench.json 2> gpurun_out/r02av_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02av_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')})
e = d['e2e']; print(e['value'], e['ms_per_step'], e['phase_ms']); print(e['set_points']); print(e['host'])
PY
tail -3 gpurun_out/r02av_bench.err
