#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/s4b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s4b_pytest_gpu.log
timeout 600 python scripts/run_all_configs.py 2> gpurun_out/s4b_all_configs.err | sed 's/CONFIG_RESULT //' > gpurun_out/s4b_all_configs.jsonl; cut -c1-330 gpurun_out/s4b_all_configs.jsonl; tail -2 gpurun_out/s4b_all_configs.err
SVMB200_PERSISTENT=0 timeout 600 python scripts/run_all_configs.py 2> /dev/null | sed 's/CONFIG_RESULT //' > gpurun_out/s4b_all_configs_twokernel.jsonl; head -1 gpurun_out/s4b_all_configs_twokernel.jsonl | cut -c1-330
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/s4b_bench_n1.json 2> gpurun_out/s4b_bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s4b_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['per_iteration_us'], d['step_parts'][0], d['e2e']['value'])
PY
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
timeout 300 python /tmp/c1.py > gpurun_out/s4b_c1_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:pg_persistent -c 1 -o gpurun_out/s4b_prof_persistent python /tmp/c1.py > gpurun_out/s4b_ncu_persistent.log 2>&1
echo "ncu persistent rc=$?"; cat gpurun_out/s4b_c1_plain.log
PROF="python bench.py --steps 1 --warmup 1 --max-iter 20 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pg_vector_kernel -s 20 -c 2 -o gpurun_out/s4b_prof_vector $PROF > gpurun_out/s4b_ncu_vector.log 2>&1; echo "ncu vector rc=$?"
