#!/bin/bash
# Round 2, session 1 (1 GPU): hardware numbers for what round 1 left untimed + the C2 isolate run.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_r2_s1.sh'
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/s1_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s1_pytest_gpu.log
timeout 300 python scripts/bench_gram.py > gpurun_out/s1_bench_gram_backoff1.log 2>&1; echo "bench_gram rc=$?"; cat gpurun_out/s1_bench_gram_backoff1.log
SVMB200_GRAM_BACKOFF=0 timeout 300 python scripts/bench_gram.py > gpurun_out/s1_bench_gram_backoff0.log 2>&1; echo "bench_gram(spin) rc=$?"; cat gpurun_out/s1_bench_gram_backoff0.log
SVMB200_GRAM_EXCLUSIVE=1 timeout 300 python scripts/bench_gram.py > gpurun_out/s1_bench_gram_exclusive.log 2>&1; echo "bench_gram(excl) rc=$?"; head -1 gpurun_out/s1_bench_gram_exclusive.log
timeout 600 python scripts/c2_isolate.py > gpurun_out/s1_c2_isolate.json 2> gpurun_out/s1_c2_isolate.err; echo "c2 rc=$?"; cat gpurun_out/s1_c2_isolate.json; tail -3 gpurun_out/s1_c2_isolate.err
timeout 600 python scripts/bench_ovr.py --classes 4 --iters 300 > gpurun_out/s1_bench_ovr_c4.jsonl 2> gpurun_out/s1_bench_ovr_c4.err; echo "bench_ovr rc=$?"; cat gpurun_out/s1_bench_ovr_c4.jsonl; tail -3 gpurun_out/s1_bench_ovr_c4.err
timeout 600 python scripts/run_all_configs.py > gpurun_out/s1_all_configs.jsonl 2> gpurun_out/s1_all_configs.err; echo "all_configs rc=$?"; cat gpurun_out/s1_all_configs.jsonl | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 2 -c 1 -o gpurun_out/s1_prof_gram python scripts/bench_gram.py > gpurun_out/s1_ncu_gram.log 2>&1
echo "ncu gram rc=$?"
nproc; free -g | head -2
