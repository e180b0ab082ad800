#!/bin/bash
# round 2, GPU session W: final kernels: GPU tests, ncu --set full of the sweeps, launch list of the bench command, sanitizers
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02w_tests.log
tail -4 gpurun_out/r02w_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_props|k_grad|k_pairs|k_reduce' -c 12 -o gpurun_out/r02w_sweeps python scripts/prof_kernels.py 8 stress 0 2 > gpurun_out/r02w_ncu.log 2>&1
ls -la gpurun_out/r02w_sweeps.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02w_launches.csv python bench.py --steps 2 --warmup 1 --no-partition --replicas 0 --no-hmc-step --no-cpu-baseline > gpurun_out/r02w_b.log 2>&1
wc -l gpurun_out/r02w_launches.csv
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/racecheck_run.py > gpurun_out/r02w_racecheck.log 2>&1; tail -4 gpurun_out/r02w_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/racecheck_run.py > gpurun_out/r02w_memcheck.log 2>&1; tail -3 gpurun_out/r02w_memcheck.log
