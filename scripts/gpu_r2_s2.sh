#!/bin/bash
# Round 2, session 2 (1 GPU): new fit path (device variance / SV gather / PDL / pipelined polls), reference arm on the box,
# K1 with the L2 store policy and integer powers, K6 at scale, ncu evidence.
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/s2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s2_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/s2_bench_n1.json 2> gpurun_out/s2_bench_n1.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/s2_bench_n1.json; tail -3 gpurun_out/s2_bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s2_bench_ref.json 2> gpurun_out/s2_bench_ref.err; echo "ref rc=$?"; cut -c1-2500 gpurun_out/s2_bench_ref.json; tail -3 gpurun_out/s2_bench_ref.err
timeout 600 python scripts/run_all_configs.py 2> gpurun_out/s2_all_configs.err | sed 's/CONFIG_RESULT //' > gpurun_out/s2_all_configs.jsonl; cut -c1-420 gpurun_out/s2_all_configs.jsonl
timeout 300 python scripts/bench_gram.py > gpurun_out/s2_bench_gram.log 2>&1; echo "bench_gram rc=$?"; cat gpurun_out/s2_bench_gram.log
timeout 600 python scripts/bench_decision.py > gpurun_out/s2_bench_decision.jsonl 2> gpurun_out/s2_bench_decision.err; echo "decision rc=$?"; cat gpurun_out/s2_bench_decision.jsonl; tail -2 gpurun_out/s2_bench_decision.err
PROF="python bench.py --steps 1 --warmup 1 --max-iter 20 --no-cpu-baseline"
timeout 300 $PROF > gpurun_out/s2_plain_prof.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches_maxiter20.csv $PROF > gpurun_out/s2_ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 2 -c 1 -o gpurun_out/s2_prof_gram python scripts/bench_gram.py > gpurun_out/s2_ncu_gram.log 2>&1
echo "ncu gram rc=$?"
