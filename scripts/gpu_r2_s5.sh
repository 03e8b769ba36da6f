#!/bin/bash
# Round 2, session 5 (8 GPUs): scaling of both launch models on C4, multi-GPU correctness incl. the back-to-back stress,
# the single-process device group, C5 against the oracle.
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29518 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/s5_bench_n${N}_torchrun.json 2> gpurun_out/s5_bench_n${N}_torchrun.err; echo "bench torchrun rc=$?"; grep '"metric"' gpurun_out/s5_bench_n${N}_torchrun.json | cut -c1-1200; grep -E "Error|Traceback" gpurun_out/s5_bench_n${N}_torchrun.err | head -3
timeout 600 python bench.py --devices all --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/s5_bench_n${N}_group.json 2> gpurun_out/s5_bench_n${N}_group.err; echo "bench group rc=$?"; cut -c1-1200 gpurun_out/s5_bench_n${N}_group.json; tail -2 gpurun_out/s5_bench_n${N}_group.err
SVMB200_CHECK_STRESS_AB=1 SVMB200_CHECK_SHARED_GRAM=1 timeout 600 $TR --master-port 29517 tests/multigpu_check.py > gpurun_out/s5_multigpu_check_n${N}.log 2>&1; echo "multigpu_check rc=$?"; grep "multigpu\|MULTIGPU" gpurun_out/s5_multigpu_check_n${N}.log | cut -c1-250
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -k single_process > gpurun_out/s5_pytest_group.log 2>&1; echo "pytest group rc=$?"; tail -4 gpurun_out/s5_pytest_group.log
timeout 900 $TR --master-port 29519 scripts/c5_check.py > gpurun_out/s5_c5_check.log 2>&1; echo "c5 rc=$?"; grep C5_RESULT gpurun_out/s5_c5_check.log | cut -c1-1500; grep -E "Error|Traceback" gpurun_out/s5_c5_check.log | head -3
