#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py > gpurun_out/multigpu_check_n$N.log 2>&1; echo "check rc=$?"; grep -E "multigpu|MULTIGPU|Error|error" gpurun_out/multigpu_check_n$N.log | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; grep '"metric"' gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
