#!/bin/bash
# round 2, GPU session BF: spread of the chains' wall clock with the pinned result blocks kept up to a byte budget
timeout 600 python scripts/chains_timing.py 8 12 2>&1 | tail -12
