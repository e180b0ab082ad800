#!/bin/bash
# round 2, GPU session BI: MeshHMC through set_points: trajectory / chain parity tests, chains, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trajectory.py tests/test_gpu_set_points.py tests/test_gpu_partition.py tests/test_checkpoint.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python scripts/chains_timing.py 8 5 2>&1 | tail -5
timeout 1500 python bench.py > gpurun_out/r02bi_bench.json 2> gpurun_out/r02bi_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02bi_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')})
e = d['e2e']; print(e['value'], e['reference_idiom']['value'], e['hmc_move'])
r = d['replicas']; print(r['moves_per_s'], r['ms_per_move_per_chain_concurrent'], r['ms_per_move_one_chain_alone'], r['lazy_repulsion']['moves_per_s'], r['lazy_repulsion']['moves_per_s_with_device_momenta'])
h = d['hmc_step_config2']; print({k: h[k] for k in h if k != 'what'}); print(d['partitioned']['public_api'])
PY
tail -3 gpurun_out/r02bi_bench.err
