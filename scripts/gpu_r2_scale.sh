#!/bin/bash
# Round 2: strong scaling of C4 on ONE 8-GPU box (the driver computes its own; this is the builder's same-box table).
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "n1 rc=$?"
for N in 2 4 8; do
  timeout 600 $TR --nproc-per-node $N --master-port $((29520 + N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; echo "n$N rc=$?"
done
timeout 600 python bench.py --devices all --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n8_group.json 2> gpurun_out/scale_n8_group.err; echo "group rc=$?"
python - <<'PY'
import json
base = None
for tag in ('n1', 'n2', 'n4', 'n8', 'n8_group'):
    try:
        d = json.loads([l for l in open(f'gpurun_out/scale_{tag}.json').read().splitlines() if l.startswith('{')][-1])
    except Exception as e:
        print(tag, 'FAILED', e); continue
    if base is None: base = d['value']
    print(tag, 'gpus', d['n_gpus'], 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'fit_ms', round(d['ms_per_step'], 1),
          'eff', round(d['value'] / base / d['n_gpus'], 3), 'per_iter', {k: round(v, 1) for k, v in d['per_iteration_us'].items()},
          'parity', d['parity']['max_abs_dalpha'], d['parity']['same_support'], 'traffic', d['roofline']['traffic'])
PY
