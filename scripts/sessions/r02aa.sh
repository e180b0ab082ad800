#!/bin/bash
# round 2, GPU session AA (2 GPUs): multi-process partition test, 2-rank bench (partitioned mesh, chains)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_partition_multiprocess.py tests/test_gpu_partition.py -m gpu -q -x 2>&1 | tail -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02aa_bench_n2.json 2> gpurun_out/r02aa_bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r02aa_bench_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02aa_bench_n2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}); print(d['e2e']['value']); print(d['replicas']['moves_per_s'], d['replicas']['lazy_repulsion']); print(d['config']['partitioned'])
PY
