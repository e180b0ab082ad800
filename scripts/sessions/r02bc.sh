#!/bin/bash
# round 2, GPU session BC: spread of the chains' wall clock over repeated rounds; this build and the last commit before this session's stream changes
timeout 600 python scripts/chains_timing.py 8 8 2>&1 | tail -9
