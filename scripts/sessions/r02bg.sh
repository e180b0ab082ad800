#!/bin/bash
# round 2, GPU session BG: final build: all GPU tests, smoke, reference arm, the default bench line, ncu --set full of the sweeps
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02bg_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02bg_tests.log
tail -4 gpurun_out/r02bg_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02bg_bench_ref.json 2> gpurun_out/r02bg_bench_ref.err; echo "ref rc=$?"
timeout 1500 python bench.py > gpurun_out/r02bg_bench.json 2> gpurun_out/r02bg_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02bg_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}); print(d['roofline']['kernel_ms'], d['roofline']['eg_frac_of_peak'], d['roofline']['leapfrog_frac_of_peak'], d['roofline']['frac'])
e = d['e2e']; print(d['leapfrog']['ms_per_step'], d['verlet_rebuild_ms'], e['value'], e['phase_ms'], e['reference_idiom']['value'], e['host'], e['hmc_move']['ms_per_move'])
r = d['replicas']; print(r['moves_per_s'], r['ms_per_move_per_chain_concurrent'], r['ms_per_move_one_chain_alone'], r['lazy_repulsion']['moves_per_s'], r['lazy_repulsion']['moves_per_s_with_device_momenta'])
h = d['hmc_step_config2']; print({k: h[k] for k in h if k != 'what'}); print(d['partitioned']['ms_per_eg'], d['partitioned']['leapfrog_ms_per_step'])
print(d['cpu_baseline']['value'], json.load(open('gpurun_out/r02bg_bench_ref.json'))['value'])
PY
tail -3 gpurun_out/r02bg_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_props|k_grad|k_pairs|k_reduce' -c 14 -o gpurun_out/r02bg_sweeps python scripts/prof_kernels.py 8 stress 0 1 2 3 > gpurun_out/r02bg_ncu.log 2>&1
ls -la gpurun_out/r02bg_sweeps.ncu-rep
