#!/bin/bash
# round 2, GPU session BA: bench line again (host-thread-heavy sections: chains, pflip), PCIe measurement moved behind them
mkdir -p gpurun_out
timeout 1500 python bench.py --no-partition > gpurun_out/r02ba_bench.json 2> gpurun_out/r02ba_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02ba_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')})
e = d['e2e']; print(e['value'], e['phase_ms'], e['reference_idiom']['value'], e['host'])
r = d['replicas']; print(r['moves_per_s'], r['ms_per_move_per_chain_concurrent'], r['ms_per_move_one_chain_alone'], r['lazy_repulsion']['moves_per_s'], r['lazy_repulsion']['moves_per_s_with_device_momenta'])
h = d['hmc_step_config2']; print({k: h[k] for k in h if k != 'what'})
PY
