#!/bin/bash
# Round 2, session 4 (1 GPU): persistent small-problem loop (C1), K1 tile bands (DRAM traffic), full gpu suite.
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/s4_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s4_pytest_gpu.log
timeout 600 python scripts/run_all_configs.py 2> gpurun_out/s4_all_configs.err | sed 's/CONFIG_RESULT //' > gpurun_out/s4_all_configs.jsonl; cut -c1-330 gpurun_out/s4_all_configs.jsonl; tail -2 gpurun_out/s4_all_configs.err
SVMB200_PERSISTENT=0 timeout 600 python scripts/run_all_configs.py 2> gpurun_out/s4_all_configs_twokernel.err | sed 's/CONFIG_RESULT //' > gpurun_out/s4_all_configs_twokernel.jsonl; head -1 gpurun_out/s4_all_configs_twokernel.jsonl | cut -c1-330
timeout 300 python scripts/bench_gram.py > gpurun_out/s4_bench_gram.log 2>&1; echo "bench_gram rc=$?"; cat gpurun_out/s4_bench_gram.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/s4_bench_n1.json 2> gpurun_out/s4_bench_n1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/s4_bench_n1.json; tail -2 gpurun_out/s4_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 2 -c 1 -o gpurun_out/s4_prof_gram python scripts/bench_gram.py > gpurun_out/s4_ncu_gram.log 2>&1
echo "ncu gram rc=$?"
cat > /tmp/c1.py <<'PY'
import sys; sys.path.insert(0, '.')
from optiml_b200.configs import make_config
from optiml_b200.ml.svm import DualSVC
from optiml_b200.ml.svm.kernels import GaussianKernel
spec, X, y = make_config('C1')
for _ in range(3):
    m = DualSVC(kernel=GaussianKernel(), C=1).fit(X, y); m.obj.release()
print(m.optimizer.iter, m.optimizer.device_ms)
PY
timeout 300 python /tmp/c1.py > gpurun_out/s4_c1_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:pg_persistent -c 1 -o gpurun_out/s4_prof_persistent python /tmp/c1.py > gpurun_out/s4_ncu_persistent.log 2>&1
echo "ncu persistent rc=$?"; cat gpurun_out/s4_c1_plain.log
