#!/bin/bash
# GPU session: parity tests, bench, ncu launch list + full capture of the streaming matvec
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
PROF="python bench.py --steps 1 --warmup 1 --max-iter 20 --no-cpu-baseline"
timeout 300 $PROF > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matvec_rows -s 30 -c 3 -o gpurun_out/prof_matvec $PROF > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 2 -c 1 -o gpurun_out/prof_gram $PROF > gpurun_out/ncu3.log 2>&1
echo "ncu gram rc=$?"
ls -la gpurun_out
