#!/bin/bash
# same-box A/B of the two peer-memory exchange protocols (flag-based build vs tagged-entry build) and NCCL
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-8}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/ab_${name}_n$N.json 2> gpurun_out/ab_${name}_n$N.err; echo "$name rc=$?"
}
run ll_1 SVMB200_EXCHANGE=p2p
run flag_1 SVMB200_EXCHANGE=p2p SVMB200_LIB=$GRAFT_REPO_ROOT/optiml_b200/_lib_flag/libsvmb200.so
run nccl_1 SVMB200_EXCHANGE=nccl
run ll_2 SVMB200_EXCHANGE=p2p
run flag_2 SVMB200_EXCHANGE=p2p SVMB200_LIB=$GRAFT_REPO_ROOT/optiml_b200/_lib_flag/libsvmb200.so
