#!/bin/bash
set -u
mkdir -p gpurun_out
for o in 512 2048; do
  echo "== planner overhead $o elements"
  (cd scripts && timeout 300 ./symv_sweep_o$o 50000 200 shard 8) > gpurun_out/sy10_sweep_shard8_o$o.log 2>&1; grep "^TR\|planned for 296" gpurun_out/sy10_sweep_shard8_o$o.log | cut -c1-250
  (cd scripts && timeout 300 ./symv_sweep_o$o 50000 200 shard 2) > gpurun_out/sy10_sweep_shard2_o$o.log 2>&1; grep "planned for 296" gpurun_out/sy10_sweep_shard2_o$o.log | cut -c1-250
done
echo "== 8192 (shipped)"
(cd scripts && timeout 300 ./symv_sweep 50000 200 shard 2) > gpurun_out/sy10_sweep_shard2_o8192.log 2>&1; grep "^TR\|planned for 296" gpurun_out/sy10_sweep_shard2_o8192.log | cut -c1-250
