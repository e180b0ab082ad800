#!/bin/bash
# round 2, GPU session J: final kernels: tests, ncu --set full, launch list, racecheck, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02j_tests.log
tail -4 gpurun_out/r02j_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_props|k_grad|k_pairs|k_reduce' -c 12 -o gpurun_out/r02_sweeps python scripts/prof_kernels.py 8 stress 0 2 > gpurun_out/r02j_ncu.log 2>&1
ls -la gpurun_out/r02_sweeps.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-partition --replicas 0 --no-hmc-step --no-cpu-baseline > gpurun_out/r02j_b.log 2>&1
wc -l gpurun_out/r02_launches.csv
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/racecheck_run.py > gpurun_out/r02_racecheck.log 2>&1; tail -6 gpurun_out/r02_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/racecheck_run.py > gpurun_out/r02_memcheck.log 2>&1; tail -4 gpurun_out/r02_memcheck.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02j_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}); print(d['roofline']['kernel_ms'], d['roofline']['eg_frac_of_peak'], d['roofline']['leapfrog_frac_of_peak']); print(d['leapfrog']['ms_per_step'], d['verlet_rebuild_ms'], d['e2e']['value'])
PY
