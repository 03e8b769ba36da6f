#!/bin/bash
# first GPU session: smoke, exploratory parity table, full-size C4
set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 900 python scripts/explore_parity.py --full > gpurun_out/explore.log 2>&1; echo "explore rc=$?"; tail -40 gpurun_out/explore.log
