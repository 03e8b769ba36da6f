#!/bin/bash
set -u
mkdir -p gpurun_out
(cd scripts && timeout 300 ./symv_sweep 50000 200 shard 8) > gpurun_out/sy9_sweep_shard8.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/sy9_sweep_shard8.log
(cd scripts && timeout 300 ./symv_sweep 50000 200 shard 4) > gpurun_out/sy9_sweep_shard4.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/sy9_sweep_shard4.log
