#!/bin/bash
# round 2, GPU session BJ (2 GPUs): bench line of the final build, partition tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_partition.py -m gpu -q -x 2>&1 | tail -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02bj_bench_n2.json 2> gpurun_out/r02bj_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
line = [l for l in open('gpurun_out/r02bj_bench_n2.json') if l.startswith('{')][-1]
d = json.loads(line)
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e']['value'], d['e2e']['phase_ms'])
print(d['leapfrog']['ms_per_step'], d['replicas']['moves_per_s'], d['replicas']['lazy_repulsion']['moves_per_s'])
p = d['partitioned']; print({k: p.get(k) for k in ('ranks', 'ms_per_eg', 'leapfrog_ms_per_step', 'verlet_rebuild_ms', 'energy_rel_diff_vs_whole_mesh_on_one_device')})
PY
tail -3 gpurun_out/r02bj_bench_n2.err
