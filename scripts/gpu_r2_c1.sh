#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pg.py tests/test_gpu_estimators.py -q -m gpu -x -k "persistent or c1_full or iteration_map or bcqp" > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c1_pytest.log
cat > /tmp/c1.py <<'PY'
import sys; sys.path.insert(0, '.')
from optiml_b200.configs import make_config
from optiml_b200.ml.svm import DualSVC
from optiml_b200.ml.svm.kernels import GaussianKernel
spec, X, y = make_config('C1')
ms = []
for _ in range(6):
    m = DualSVC(kernel=GaussianKernel(), C=1).fit(X, y); m.obj.release(); ms.append(round(m.optimizer.device_ms, 3))
print('C1 device ms per 1000 iterations:', ms, 'best it/s', round(1e6 / min(ms)))
PY
timeout 300 python /tmp/c1.py > gpurun_out/c1_plain.log 2>&1; cat gpurun_out/c1_plain.log
SVMB200_PERSISTENT=0 timeout 300 python /tmp/c1.py 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pg_persistent -c 1 -o gpurun_out/c1_prof_persistent python /tmp/c1.py > gpurun_out/c1_ncu.log 2>&1; echo "ncu rc=$?"
