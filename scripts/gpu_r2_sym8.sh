#!/bin/bash
# round 2: the graded plan (short bands, cut panels) on 8 GPUs: symmetric cases of multigpu_check, bench --symmetric at 8 and 4 GPUs
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
SVMB200_CHECK_DEFAULT=0 SVMB200_CHECK_STRESS=0 timeout 400 $TR --nproc-per-node=$N --master-port 29581 tests/multigpu_check.py > gpurun_out/sy8_multigpu_check_symmetric_n$N.log 2>&1
echo "multigpu_check rc=$?"; grep "symmetric\|MULTIGPU" gpurun_out/sy8_multigpu_check_symmetric_n$N.log | tail -n 8 | cut -c1-260
for G in 8 4; do
  [ $G -gt $N ] && continue
  DEV=$(seq -s, 0 $((G-1)))
  CUDA_VISIBLE_DEVICES=$DEV timeout 400 $TR --nproc-per-node=$G --master-port $((29590+G)) bench.py --gpus $G --steps 5 --warmup 3 --symmetric > gpurun_out/sy8_bench_n${G}_symmetric.json 2> gpurun_out/sy8_bench_n${G}_symmetric.err
  echo "bench symmetric N=$G rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/sy8_bench_n${G}_symmetric.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','fit_s','per_iteration_us')}, 'e2e', d['e2e']['value'], 'parity', d['parity']['meets_north_star'], d['parity']['max_abs_dalpha'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'])"
done
