#!/bin/bash
# round 2, GPU session AT: split pair sweep build: GPU tests, smoke, bench line, reference arm, launch list, sanitizers
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02at_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02at_tests.log
tail -3 gpurun_out/r02at_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r02at_bench.json 2> gpurun_out/r02at_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02at_bench_ref.json 2> gpurun_out/r02at_bench_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/r02at_bench_ref.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02at_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}); print(d['roofline']['kernel_ms'], d['roofline']['eg_frac_of_peak'], d['roofline']['leapfrog_frac_of_peak'], d['roofline']['frac'])
print(d['leapfrog']['ms_per_step'], d['verlet_rebuild_ms'], d['e2e']['value'], d['e2e']['phase_ms'], d['e2e']['hmc_move']['ms_per_move'], d['config']['host_cores_bound_per_rank'])
print(d['replicas']['moves_per_s'], d['replicas']['lazy_repulsion']); print(d['hmc_step_config2']['hmc_move_ms'], d['hmc_step_config2']['pflip_ms']); print(d['partitioned']['ms_per_eg'], d['partitioned']['leapfrog_ms_per_step'], d['partitioned']['reference_golden'])
print(d['cpu_baseline'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02at_launches.csv python bench.py --steps 2 --warmup 1 --no-partition --replicas 0 --no-hmc-step --no-cpu-baseline > gpurun_out/r02at_b.log 2>&1
wc -l gpurun_out/r02at_launches.csv
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/racecheck_run.py split > gpurun_out/r02at_memcheck.log 2>&1; tail -3 gpurun_out/r02at_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/racecheck_run.py > gpurun_out/r02at_racecheck.log 2>&1; tail -3 gpurun_out/r02at_racecheck.log
