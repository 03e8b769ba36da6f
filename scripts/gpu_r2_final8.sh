#!/bin/bash
# round 2, last 8-GPU session: the symmetric cases of multigpu_check (expectation fixed), the default bench line at 8 GPUs
# (which now carries the symmetric leg), C5 with the symmetric pass
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
SVMB200_CHECK_DEFAULT=0 SVMB200_CHECK_STRESS=0 timeout 400 $TR --master-port 29571 tests/multigpu_check.py > gpurun_out/fin8_multigpu_check_symmetric_n$N.log 2>&1
echo "multigpu_check rc=$?"; grep "symmetric\|MULTIGPU" gpurun_out/fin8_multigpu_check_symmetric_n$N.log | tail -n 8
timeout 400 $TR --master-port 29572 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/fin8_bench_n$N.json 2> gpurun_out/fin8_bench_n$N.err
echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/fin8_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','fit_s','per_iteration_us','product_pass')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'])
print('symmetric leg', json.dumps(d.get('symmetric_pass'))[:1500])"
SVMB200_SYMMETRIC=1 timeout 400 $TR --master-port 29573 scripts/c5_check.py > gpurun_out/fin8_c5_check_symmetric.log 2>&1
echo "c5 symmetric rc=$?"; grep C5_RESULT gpurun_out/fin8_c5_check_symmetric.log | cut -c1-1200; grep -E "Error|Traceback" gpurun_out/fin8_c5_check_symmetric.log | head -3
