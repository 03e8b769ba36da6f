#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== shipped"; (cd scripts && timeout 300 ./symv_sweep 50000 100 shard 8) > gpurun_out/sy11_shipped.log 2>&1; grep "^TR\|planned for 296" gpurun_out/sy11_shipped.log | cut -c1-250
echo "== u hoisted"; (cd scripts && timeout 300 ./symv_sweep_try 50000 100 shard 8) > gpurun_out/sy11_try.log 2>&1; grep "^TR\|planned for 296" gpurun_out/sy11_try.log | cut -c1-250
echo "== shipped again"; (cd scripts && timeout 300 ./symv_sweep 50000 100 one) > gpurun_out/sy11_shipped2.log 2>&1; grep "^TR" gpurun_out/sy11_shipped2.log | cut -c1-250
echo "== u hoisted again"; (cd scripts && timeout 300 ./symv_sweep_try 50000 100 one) > gpurun_out/sy11_try2.log 2>&1; grep "^TR" gpurun_out/sy11_try2.log | cut -c1-250
