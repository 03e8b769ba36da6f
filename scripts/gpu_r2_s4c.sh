#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/s4c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s4c_pytest_gpu.log
timeout 600 python scripts/run_all_configs.py 2> gpurun_out/s4c_all_configs.err | sed 's/CONFIG_RESULT //' > gpurun_out/s4c_all_configs.jsonl; cut -c1-330 gpurun_out/s4c_all_configs.jsonl; tail -2 gpurun_out/s4c_all_configs.err
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
timeout 300 python /tmp/c1.py > gpurun_out/s4c_c1_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:pg_persistent -c 1 -o gpurun_out/s4c_prof_persistent python /tmp/c1.py > gpurun_out/s4c_ncu_persistent.log 2>&1
echo "ncu persistent rc=$?"; cat gpurun_out/s4c_c1_plain.log
