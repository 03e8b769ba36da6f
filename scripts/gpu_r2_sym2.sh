#!/bin/bash
set -u
mkdir -p gpurun_out
(cd scripts && timeout 600 ./symv_sweep 50000 60) > gpurun_out/sy3_sweep.log 2>&1
echo "sweep rc=$?"
tail -n 16 gpurun_out/sy3_sweep.log | cut -c1-330
timeout 900 python -m pytest tests/test_gpu_symmetric.py -x -q > gpurun_out/sy3_pytest_sym.log 2>&1
echo "pytest sym rc=$?"
tail -n 5 gpurun_out/sy3_pytest_sym.log
