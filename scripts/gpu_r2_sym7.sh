#!/bin/bash
set -u
mkdir -p gpurun_out
(cd scripts && timeout 300 ./symv_sweep 50000 40 one) > gpurun_out/sy7_sweep_one.log 2>&1; echo "sweep rc=$?"; tail -n 2 gpurun_out/sy7_sweep_one.log | cut -c1-330
timeout 600 python -m pytest tests/test_gpu_symmetric.py -x -q > gpurun_out/sy7_pytest_sym.log 2>&1; echo "pytest sym rc=$?"; tail -n 3 gpurun_out/sy7_pytest_sym.log
timeout 400 python bench.py --steps 3 --warmup 2 --symmetric --no-cpu-baseline > gpurun_out/sy7_bench_n1_symmetric.json 2> gpurun_out/sy7_bench_n1_symmetric.err
echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/sy7_bench_n1_symmetric.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','fit_s','per_iteration_us')}, 'e2e', d['e2e']['value'], 'parity', d['parity']['meets_north_star'], d['parity']['max_abs_dalpha'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'])"
