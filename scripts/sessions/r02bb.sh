#!/bin/bash
# round 2, GPU session BB: the chains section with and without stream priorities, twice each
mkdir -p gpurun_out
for rep in 1 2; do for prio in 1 0; do
TM_STREAM_PRIO=$prio timeout 900 python bench.py --no-partition --no-hmc-step --no-cpu-baseline --steps 10 > gpurun_out/r02bb_bench_p${prio}_${rep}.json 2> gpurun_out/r02bb.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02bb_bench_p${prio}_${rep}.json'))
r = d['replicas']; print('prio', $prio, 'rep', $rep, d['value'], d['e2e']['value'], round(r['moves_per_s']), r['ms_per_move_per_chain_concurrent'], r['ms_per_move_one_chain_alone'], round(r['lazy_repulsion']['moves_per_s']), round(r['lazy_repulsion']['moves_per_s_with_device_momenta']))
PY
done; done
