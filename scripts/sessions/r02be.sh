#!/bin/bash
# round 2, GPU session BE: spread of the chains' wall clock against the interpreter's thread switch interval
for si in 0.005 0.0002 0.00002; do echo "== switch interval $si"; timeout 600 python scripts/chains_timing.py 8 8 $si 2>&1 | tail -8; done
