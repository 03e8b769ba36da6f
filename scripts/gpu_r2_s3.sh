#!/bin/bash
# Round 2, session 3 (2 GPUs): multi-GPU correctness (torchrun ranks + stress with the creation barrier A/B), the
# single-process device group, and both launch models timed on C4.
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
SVMB200_CHECK_STRESS_AB=1 SVMB200_CHECK_SHARED_GRAM=1 timeout 900 $TR --master-port 29517 tests/multigpu_check.py > gpurun_out/s3_multigpu_check_n2.log 2>&1; echo "multigpu_check rc=$?"; grep "multigpu\|MULTIGPU" gpurun_out/s3_multigpu_check_n2.log | cut -c1-260
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -k single_process > gpurun_out/s3_pytest_group.log 2>&1; echo "pytest group rc=$?"; tail -5 gpurun_out/s3_pytest_group.log
timeout 600 $TR --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/s3_bench_n2_torchrun.json 2> gpurun_out/s3_bench_n2_torchrun.err; echo "bench torchrun rc=$?"; cut -c1-1800 gpurun_out/s3_bench_n2_torchrun.json; tail -2 gpurun_out/s3_bench_n2_torchrun.err
timeout 600 python bench.py --devices 0,1 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/s3_bench_n2_group.json 2> gpurun_out/s3_bench_n2_group.err; echo "bench group rc=$?"; cut -c1-1800 gpurun_out/s3_bench_n2_group.json; tail -2 gpurun_out/s3_bench_n2_group.err
SVMB200_MATVEC_L2_HINT=1 timeout 600 $TR --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/s3_bench_n2_torchrun_l2hint.json 2> gpurun_out/s3_bench_n2_torchrun_l2hint.err; echo "bench torchrun hint rc=$?"; cut -c1-400 gpurun_out/s3_bench_n2_torchrun_l2hint.json
