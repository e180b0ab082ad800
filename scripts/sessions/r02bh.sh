#!/bin/bash
# round 2, GPU session BH: chains with size classes in the device block cache
timeout 600 python scripts/chains_timing.py 8 8 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_nlist.py tests/test_gpu_flips.py -m gpu -q -x 2>&1 | tail -2
