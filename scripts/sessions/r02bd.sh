#!/bin/bash
# round 2, GPU session BD: spread of the chains' wall clock: the commit before this session's stream changes (_old worktree), then this build
echo "== old (dedec01)"; (cd _old && timeout 600 python scripts/chains_timing.py 8 8 2>&1 | tail -8)
echo "== new"; timeout 600 python scripts/chains_timing.py 8 8 2>&1 | tail -8
