#!/bin/bash
# round 2, symmetric pass on row blocks: first hardware session on N GPUs (N = number visible)
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 600 python -m pytest tests/test_gpu_symmetric.py -x -q > gpurun_out/sy5_pytest_sym.log 2>&1
echo "pytest sym rc=$?"; tail -n 3 gpurun_out/sy5_pytest_sym.log
SVMB200_CHECK_STRESS=60 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29531 \
    tests/multigpu_check.py > gpurun_out/sy5_multigpu_check_n$N.log 2>&1
echo "multigpu_check rc=$?"; grep "multigpu\|MULTIGPU" gpurun_out/sy5_multigpu_check_n$N.log | tail -n 20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/sy5_bench_n${N}_full.json 2> gpurun_out/sy5_bench_n${N}_full.err
echo "bench full rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/sy5_bench_n${N}_full.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','fit_s','per_iteration_us','parity')}, d['e2e']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 3 --warmup 2 --symmetric > gpurun_out/sy5_bench_n${N}_symmetric.json 2> gpurun_out/sy5_bench_n${N}_symmetric.err
echo "bench symmetric rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/sy5_bench_n${N}_symmetric.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','fit_s','per_iteration_us','parity','product_pass')}, d['e2e']['value'])"
tail -n 5 gpurun_out/sy5_bench_n${N}_symmetric.err
timeout 600 python bench.py --devices all --steps 3 --warmup 2 --symmetric --no-cpu-baseline > gpurun_out/sy5_bench_n${N}_group_symmetric.json 2> gpurun_out/sy5_bench_n${N}_group_symmetric.err
echo "bench group symmetric rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/sy5_bench_n${N}_group_symmetric.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','fit_s','parity','product_pass','launch_model')}, d['e2e']['value'])"
tail -n 5 gpurun_out/sy5_bench_n${N}_group_symmetric.err
