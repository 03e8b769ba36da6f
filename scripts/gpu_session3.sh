#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
from optiml_b200.configs import make_config
from optiml_b200.ml.svm import DualSVC
from optiml_b200.ml.svm.kernels import GaussianKernel
spec, X, y = make_config('C4')
for i in range(3):
    t = time.perf_counter(); m = DualSVC(kernel=GaussianKernel(), C=1).fit(X, y); dt = time.perf_counter() - t
    print('fit', dt, m.fit_times_, 'pg device ms', m.optimizer.device_ms)
    t = time.perf_counter(); m.obj.release(); print('release', time.perf_counter() - t)
PY
