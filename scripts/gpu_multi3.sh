#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-8}; EX=${2:-p2p}
export SVMB200_EXCHANGE=$EX
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py > gpurun_out/multigpu_check_n${N}_$EX.log 2>&1; echo "check $EX rc=$?"; grep -E "multigpu|MULTIGPU|Error|error|Traceback" gpurun_out/multigpu_check_n${N}_$EX.log | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_$EX.json 2> gpurun_out/bench_n${N}_$EX.err; echo "bench $EX rc=$?"; grep '"metric"' gpurun_out/bench_n${N}_$EX.json; grep -E "Error|error|Traceback" gpurun_out/bench_n${N}_$EX.err | head -5
