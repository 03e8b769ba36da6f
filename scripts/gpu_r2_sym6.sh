#!/bin/bash
# round 2, symmetric pass on row blocks at 8 GPUs: correctness (multigpu_check) and the same-box scaling table
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
SVMB200_CHECK_STRESS=60 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541 \
    tests/multigpu_check.py > gpurun_out/sy6_multigpu_check_n$N.log 2>&1
echo "multigpu_check rc=$?"; grep "symmetric\|MULTIGPU\|stress" gpurun_out/sy6_multigpu_check_n$N.log | tail -n 12
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$1', {k:d.get(k) for k in ('value','n_gpus','fit_s','per_iteration_us','product_pass')}, 'e2e', d['e2e']['value'], 'parity', (d.get('parity') or {}).get('meets_north_star'), (d.get('parity') or {}).get('max_abs_dalpha'))"; }
for G in 8 4 2; do
  [ $G -gt $N ] && continue
  DEV=$(seq -s, 0 $((G-1)))
  CUDA_VISIBLE_DEVICES=$DEV timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1 --master-port $((29550+G)) \
      bench.py --gpus $G --steps 4 --warmup 2 --symmetric > gpurun_out/sy6_bench_n${G}_symmetric.json 2> gpurun_out/sy6_bench_n${G}_symmetric.err
  echo "bench symmetric N=$G rc=$?"; show gpurun_out/sy6_bench_n${G}_symmetric.json
done
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --steps 3 --warmup 2 --symmetric --no-cpu-baseline > gpurun_out/sy6_bench_n1_symmetric.json 2> gpurun_out/sy6_bench_n1_symmetric.err
echo "bench symmetric N=1 rc=$?"; show gpurun_out/sy6_bench_n1_symmetric.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus $N --steps 4 --warmup 2 > gpurun_out/sy6_bench_n${N}_full.json 2> gpurun_out/sy6_bench_n${N}_full.err
echo "bench full N=$N rc=$?"; show gpurun_out/sy6_bench_n${N}_full.json
