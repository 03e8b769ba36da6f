#!/bin/bash
# round 2, last 1-GPU session: gpu suite, default bench line (with the symmetric leg and the reference arm), launch list of the
# bench command in symmetric mode, C3 A/B, smoke
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/fin_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/fin_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/fin_bench_n1.json 2> gpurun_out/fin_bench_n1.err; echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/fin_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','fit_s','per_iteration_us','product_pass','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'])
print('cpu', json.dumps(d.get('cpu_baseline'))[:400])
print('symmetric leg', json.dumps(d.get('symmetric_pass'))[:1800])"; tail -2 gpurun_out/fin_bench_n1.err
PROF="python bench.py --steps 1 --warmup 1 --max-iter 20 --no-cpu-baseline --symmetric"
timeout 300 $PROF > gpurun_out/fin_plain_prof_symmetric.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fin_launches_symmetric_maxiter20.csv $PROF > gpurun_out/fin_ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 600 python scripts/bench_symmetric.py --config C3 --steps 2 --warmup 1 --max-iter 300 > gpurun_out/fin_ab_c3.jsonl 2> gpurun_out/fin_ab_c3.err; echo "ab c3 rc=$?"; cut -c1-600 gpurun_out/fin_ab_c3.jsonl
python __graft_entry__.py smoke 2>&1 | tail -2
