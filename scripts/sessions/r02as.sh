#!/bin/bash
# round 2, GPU session AS: leapfrog trajectory, graph replay vs direct launches, split vs joined
mkdir -p gpurun_out
for env in "TM_PAIR_SPLIT=2 TM_VV_GRAPH=1" "TM_PAIR_SPLIT=2 TM_VV_GRAPH=0" "TM_PAIR_SPLIT=0 TM_VV_GRAPH=1" "TM_PAIR_SPLIT=0 TM_VV_GRAPH=0"; do
  echo "== $env"
  env $env timeout 300 python scripts/vv_timing.py 8 10 2>&1 | tail -4
done
