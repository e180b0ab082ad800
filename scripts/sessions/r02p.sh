#!/bin/bash
# round 2, GPU session P: GPU tests (checkpoint hand-off, PLY, flip_with), kernel timings incl. the composites
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02p_tests.log
tail -25 gpurun_out/r02p_tests.log
timeout 300 python scripts/kernel_bench.py 8 stress verlet-list > gpurun_out/r02p_kernels.jsonl 2>gpurun_out/r02p_kb.err; cat gpurun_out/r02p_kernels.jsonl
