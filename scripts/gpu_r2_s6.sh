#!/bin/bash
# Round 2, session 6 (1 GPU): final-tree evidence -- gpu suite, K2 traffic per GPU count (ncu) and shard timing, default bench,
# launch list of the bench command, K2 / K3 full captures.
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/s6_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s6_pytest_gpu.log
timeout 300 python scripts/matvec_traffic.py time > gpurun_out/s6_matvec_shard_timing.jsonl 2>&1; echo "shard timing rc=$?"; cat gpurun_out/s6_matvec_shard_timing.jsonl
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:matvec_seg_kernel --csv --log-file gpurun_out/s6_matvec_traffic.csv python scripts/matvec_traffic.py > gpurun_out/s6_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
timeout 900 python bench.py > gpurun_out/s6_bench_n1.json 2> gpurun_out/s6_bench_n1.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/s6_bench_n1.json; tail -2 gpurun_out/s6_bench_n1.err
PROF="python bench.py --steps 1 --warmup 1 --max-iter 20 --no-cpu-baseline"
timeout 300 $PROF > gpurun_out/s6_plain_prof.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s6_launches_maxiter20.csv $PROF > gpurun_out/s6_ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:matvec_seg_kernel -s 10 -c 3 -o gpurun_out/s6_prof_matvec $PROF > gpurun_out/s6_ncu2.log 2>&1; echo "ncu matvec rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pg_vector_kernel -s 20 -c 2 -o gpurun_out/s6_prof_vector $PROF > gpurun_out/s6_ncu3.log 2>&1; echo "ncu vector rc=$?"
python __graft_entry__.py smoke 2>&1 | tail -2
