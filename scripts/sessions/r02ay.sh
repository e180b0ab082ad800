#!/bin/bash
# round 2, GPU session AY (8 GPUs): the bench line at 8 ranks (replicas, partitioned 10,485,762-vertex mesh, chains) on the split / set_points build
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02ay_bench_n8.json 2> gpurun_out/r02ay_bench_n8.err; echo "bench rc=$?"; tail -3 gpurun_out/r02ay_bench_n8.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02ay_bench_n8.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}); print(d['e2e']['value'], d['e2e']['phase_ms'], d['e2e']['reference_idiom']['value'], d['e2e']['reference_idiom']['phase_ms']); print(d['replicas']['moves_per_s'], d['replicas']['lazy_repulsion']['moves_per_s'], d['replicas']['lazy_repulsion']['moves_per_s_with_device_momenta']); print(d['config']['partitioned'])
PY
