#!/bin/bash
# round 2, last sessions on the final tree
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/fin2_pytest_gpu_n$N.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/fin2_pytest_gpu_n$N.log
if [ $N -eq 1 ]; then
  timeout 600 python bench.py --no-cpu-baseline > gpurun_out/fin2_bench_n1.json 2> gpurun_out/fin2_bench_n1.err; echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/fin2_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','fit_s','per_iteration_us','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'])
sp=d['symmetric_pass']; print('symmetric', sp['value'], sp['e2e']['value'], sp['per_iteration_us'], sp['roofline']['achieved'], sp['roofline']['frac'], sp['parity']['meets_north_star'], sp['parity']['max_abs_dalpha'])"
  python __graft_entry__.py smoke 2>&1 | tail -2
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/fin2_bench_n$N.json 2> gpurun_out/fin2_bench_n$N.err; echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/fin2_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','fit_s','per_iteration_us','gpu_launches')}, 'e2e', d['e2e']['value'])
sp=d['symmetric_pass']; print('symmetric', sp['value'], sp['e2e']['value'], sp['per_iteration_us'], sp['parity']['meets_north_star'])"
fi
