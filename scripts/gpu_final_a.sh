#!/bin/bash
# final single-box session: tests (incl. 2-GPU torchrun test), benches at N=1,2, reference arm, ncu evidence
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
PROF="python bench.py --steps 1 --warmup 1 --max-iter 20 --no-cpu-baseline"
timeout 300 $PROF > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matvec_seg -s 30 -c 3 -o gpurun_out/prof_matvec $PROF > gpurun_out/ncu2.log 2>&1; echo "ncu matvec rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pg_vector_kernel -s 30 -c 2 -o gpurun_out/prof_vector $PROF > gpurun_out/ncu4.log 2>&1; echo "ncu vec rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 2 -c 1 -o gpurun_out/prof_gram $PROF > gpurun_out/ncu3.log 2>&1; echo "ncu gram rc=$?"
python scripts/bench_gram.py > gpurun_out/bench_gram.log 2>&1; tail -6 gpurun_out/bench_gram.log
