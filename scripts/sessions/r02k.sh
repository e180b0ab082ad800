#!/bin/bash
# round 2, GPU session K (2 GPUs): k_grad capture, 2-rank bench (partitioned mesh, chains), multi-process partition test
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_grad|k_props' -c 8 --launch-skip 6 -o gpurun_out/r02_grad python scripts/prof_kernels.py 8 stress 2 1 > gpurun_out/r02k_ncu.log 2>&1
ls -la gpurun_out/r02_grad.ncu-rep
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02k_bench_n2.json 2> gpurun_out/r02k_bench_n2.err; echo "bench rc=$?"; tail -5 gpurun_out/r02k_bench_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02k_bench_n2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}); print(d['e2e']); print(d['replicas']); print(d['config']['partitioned']); print(d['partitioned'])
PY
