#!/bin/bash
set -u
mkdir -p gpurun_out
cd scripts
timeout 300 ./symv_sweep 50000 20 one > ../gpurun_out/sy4_plain.log 2>&1
echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:symv_tile -s 4 -c 2 -o ../gpurun_out/sy4_symv_tile ./symv_sweep 50000 4 one > ../gpurun_out/sy4_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 ../gpurun_out/sy4_ncu.log
timeout 600 ncu --set full --clock-control none -k regex:symv_combine -s 4 -c 1 -o ../gpurun_out/sy4_symv_combine ./symv_sweep 50000 4 one > ../gpurun_out/sy4_ncu2.log 2>&1
echo "ncu2 rc=$?"
