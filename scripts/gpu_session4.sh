#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
bash scripts/gpu_multi.sh 2
PROF="python bench.py --steps 1 --warmup 1 --max-iter 20 --no-cpu-baseline"
timeout 300 $PROF > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matvec_seg -s 30 -c 3 -o gpurun_out/prof_matvec $PROF > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pg_vector_kernel -s 30 -c 2 -o gpurun_out/prof_vector $PROF > gpurun_out/ncu4.log 2>&1
echo "ncu vec rc=$?"
