#!/bin/bash
# round 2, GPU session BM (8 GPUs): partitioned 10,485,762-vertex mesh with the split pair sweep on every rank
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --replicas 0 --no-hmc-step > gpurun_out/r02bm_bench_n8.json 2> gpurun_out/r02bm_bench_n8.err; echo "bench rc=$?"; tail -2 gpurun_out/r02bm_bench_n8.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02bm_bench_n8.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e']['value']); print(d['partitioned'])
PY
