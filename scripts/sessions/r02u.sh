#!/bin/bash
# round 2, GPU session U: flip sweeps: exclusion sets, block-wide Metropolis scan, patch lists: parity, timing, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flips.py tests/test_gpu_flip_statistics.py tests/test_checkpoint.py tests/test_gpu_partition.py -m gpu -q -x > gpurun_out/r02u_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02u_tests.log
tail -6 gpurun_out/r02u_tests.log
timeout 300 python scripts/prof_pflip.py 2>&1 | tail -6
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 1100 --csv --log-file gpurun_out/r02u_pflip_launches.csv python scripts/prof_pflip.py > gpurun_out/r02u_pflip.log 2>&1
python scripts/ncu_launch_agg.py gpurun_out/r02u_pflip_launches.csv | head -8
