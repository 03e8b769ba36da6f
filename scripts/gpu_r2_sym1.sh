#!/bin/bash
# round 2, symmetric pass, first hardware session: shape sweep of K2s, its gpu tests, the suites that cover the K3 operand
# prefetch, and the C4 A/B through the public API
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/sy1_gpu.txt 2>&1
(cd scripts && timeout 600 ./symv_sweep 50000 60) > gpurun_out/sy1_sweep.log 2>&1
echo "sweep rc=$?"
tail -n 16 gpurun_out/sy1_sweep.log
timeout 900 python -m pytest tests/test_gpu_symmetric.py -x -q > gpurun_out/sy1_pytest_sym.log 2>&1
echo "pytest sym rc=$?"
tail -n 5 gpurun_out/sy1_pytest_sym.log
timeout 900 python -m pytest tests/test_gpu_pg.py tests/test_gpu_estimators.py tests/test_gpu_frank_wolfe.py -x -q > gpurun_out/sy1_pytest_pg.log 2>&1
echo "pytest pg rc=$?"
tail -n 5 gpurun_out/sy1_pytest_pg.log
timeout 600 python scripts/bench_symmetric.py --config C4 --steps 3 --warmup 1 > gpurun_out/sy1_ab_c4.jsonl 2> gpurun_out/sy1_ab_c4.err
echo "ab rc=$?"
cat gpurun_out/sy1_ab_c4.jsonl
