#!/bin/bash
# First GPU session of round 2: hardware confirmation and measurement of widening 4 (shared-Gram one-vs-rest),
# which round 1 could only verify on the host emulation.  One gpurun call, 1 GPU, ~25 min:
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_round2_first.sh'
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
# 1. the new row first (xfail non-strict: look for XPASS), then the whole gpu suite
timeout 900 python -m pytest tests/test_gpu_shared_gram.py -q -m gpu -rxX > gpurun_out/pytest_shared_gram.log 2>&1; echo "shared-gram rc=$?"; tail -15 gpurun_out/pytest_shared_gram.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
# 2. headline bench (contract line)
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
# 3. one-vs-rest at C4 size: shared Gram + lockstep vs clone-per-class
timeout 900 python scripts/bench_ovr.py --classes 4 --iters 300 > gpurun_out/bench_ovr_c4.jsonl 2> gpurun_out/bench_ovr_c4.err; echo "bench_ovr rc=$?"; cat gpurun_out/bench_ovr_c4.jsonl; tail -3 gpurun_out/bench_ovr_c4.err
# 4. shape sweep of K2 x NB
timeout 900 python scripts/sweep_multi.py > gpurun_out/sweep_multi.jsonl 2> gpurun_out/sweep_multi.err; echo "sweep rc=$?"; cat gpurun_out/sweep_multi.jsonl; tail -3 gpurun_out/sweep_multi.err
# 4b. Gram kernel: interior-tile epilogue (new) and producer back-off A/B (round-1 reference: C4 Gaussian 26.6 ms)
timeout 600 python scripts/bench_gram.py > gpurun_out/bench_gram_backoff1.log 2>&1; echo "bench_gram rc=$?"; cat gpurun_out/bench_gram_backoff1.log
SVMB200_GRAM_BACKOFF=0 timeout 600 python scripts/bench_gram.py > gpurun_out/bench_gram_backoff0.log 2>&1; echo "bench_gram (spin) rc=$?"; cat gpurun_out/bench_gram_backoff0.log
# 5. ncu: launch list of a short one-vs-rest fit, full capture of the multi-vector pass
PROF="python scripts/bench_ovr.py --classes 4 --iters 20 --skip-cloned"
timeout 300 $PROF > gpurun_out/plain_ovr.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_ovr.csv $PROF > gpurun_out/ncu_ovr1.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matvec_seg_multi -s 10 -c 3 -o gpurun_out/prof_matvec_multi $PROF > gpurun_out/ncu_ovr2.log 2>&1
echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 2 -c 1 -o gpurun_out/prof_gram_r2 python scripts/bench_gram.py > gpurun_out/ncu_gram_r2.log 2>&1
echo "ncu gram rc=$?"
