"""The driver's entry points on the host emulation of the kernels: `__graft_entry__.smoke()` as is, and `bench.py`
on a shrunken workload -- the JSON line must carry every key of the measurement contract (numbers are meaningless
here: the "device" is a CPU).  Catches a refactor that breaks the round-end run before a GPU is involved."""
import json
import sys

import pytest

from emu import emulated_device


def test_smoke_entry_point_passes_on_the_emulation(capsys):
    import __graft_entry__ as entry
    with emulated_device() as lib:
        entry.smoke()   # one small DualSVC fit checked against the oracle: alphas, support set, decision values
        assert lib.emu_sticky_error() == 0
    assert 'smoke: n=512' in capsys.readouterr().out


def test_bench_line_carries_the_whole_contract(capsys, monkeypatch):
    import bench
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--n', '192', '--max-iter', '8', '--steps', '2', '--warmup', '1',
                                      '--no-cpu-baseline'])
    with emulated_device() as lib:
        bench.main()
        assert lib.emu_sticky_error() == 0
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith('{')]
    assert len(lines) == 1                                   # ONE JSON line
    out = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'roofline', 'e2e', 'gpu_launches', 'clocks'):
        assert key in out, key
    assert out['n_gpus'] == 1 and out['steps'] == 2 and out['warmup'] == 1 and out['dtype'] == 'f64'
    assert out['data'] == 'synthetic' and out['higher_is_better'] is True and out['vs_baseline'] is None
    assert 'workload' in out['config'] and 'model' not in out['config']
    roof = out['roofline']
    assert roof['bound'] == 'hbm' and roof['unit'] == 'GB/s' and roof['peak'] > 0
    assert roof['frac'] == pytest.approx(roof['achieved'] / roof['peak'])
    assert roof['bytes_per_launch'] == 8.0 * 192 * 192        # algorithmic bytes: one pass over the n x n matrix
    e2e = out['e2e']
    assert e2e['unit'] == out['unit'] and e2e['h2d_bytes_per_step'] > 0 and e2e['d2h_bytes_per_step'] > 0
    assert e2e['value'] != out['value']                       # measured separately, through the public API
    # 8 iterations: 1 + 8 passes and 1 + 8 (+ final state) vector launches per fit, plus Gram / norms / K5
    assert out['gpu_launches'] >= 2 * 8
    assert set(out['clocks']) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}


def test_product_refuses_to_load_the_emulated_library():
    """SVMB200_LIB selects a build variant (tuning sweeps); pointing it at the emulation must not give a CPU path"""
    import os
    import subprocess
    import emu
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = emu._builder().build()
    code = ('from optiml_b200 import _native as N\n'
            'try:\n    N.load_library()\n    print("LOADED")\n'
            'except N.NativeError as e:\n    print("REFUSED", e)\n')
    out = subprocess.run([sys.executable, '-c', code], cwd=root, env=dict(os.environ, SVMB200_LIB=lib),
                         capture_output=True, text=True, timeout=120)
    assert 'REFUSED' in out.stdout and 'no CPU fallback' in out.stdout, out.stdout + out.stderr


def test_integration_md_stubs_run_against_the_emulated_library():
    """INTEGRATION.md's ctypes stubs, executed: the host-pointer kernel / solver entry points and the device-resident
    one-vs-rest recipe, against the oracle (the `-m "not gpu"` doc test only compiles them)."""
    import os
    import re
    import numpy as np
    import emu
    from oracle import svm_oracle as O
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, 'INTEGRATION.md')) as fh:
        blocks = re.findall(r'```python\n(.*?)```', fh.read(), flags=re.S)
    assert len(blocks) >= 2
    ns = {}
    exec(compile('\n'.join(blocks).replace("'libsvmb200.so'", repr(emu._builder().build())), 'INTEGRATION.md', 'exec'), ns)
    rng = np.random.default_rng(1)
    X = rng.standard_normal((90, 6))
    gamma = 1.0 / (6 * X.var())
    K = ns['kernel_matrix'](X, None, 2, gamma)
    assert np.abs(K - O.gaussian_kernel(X)).max() <= 1e-12
    labels = rng.integers(0, 3, 90)
    signs = [np.where(labels == c, 1.0, -1.0) for c in range(3)]
    alphas, iters, statuses = ns['one_vs_rest_alphas'](X, signs, gamma, max_iter=25)
    for c in range(3):
        want = O.svc_dual_fit(X, (labels == c).astype(int), kind='gaussian', max_iter=25)
        assert iters[c] == want.pg.iter and np.abs(alphas[c] - want.alphas_).max() <= 1e-10
    Q = O.svc_dual_fit(X, (labels == 0).astype(int), kind='gaussian', max_iter=1).Q
    x, g, hist, it, status = ns['projected_gradient'](Q, -np.ones(90), np.ones(90), max_iter=25)
    want = O.projected_gradient(Q, -np.ones(90), np.ones(90), max_iter=25)
    assert it == want.iter and status == want.status and np.abs(x - want.x).max() <= 1e-10


def test_reference_arm_runs_the_unmodified_reference_on_the_same_config(capsys, monkeypatch):
    """`bench.py --impl reference`: the UNMODIFIED reference (baseline/_ref, or the container's checkout) on the full
    workload -- same `config` object as the GPU arm, `kind: "reference"`; when the host cannot hold the reference's
    matrices it says so and falls back to the NumPy port on a row sample (`kind: "port"`)."""
    import bench
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip('no reference copy (baseline/_ref or /root/reference)')
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--impl', 'reference', '--n', '700', '--max-iter', '1000', '--steps', '2', '--warmup', '1'])
    bench.main()
    line = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith('{')][-1])
    assert line['impl'] == 'reference' and line['cpu_baseline']['kind'] == 'reference' and line['value'] > 0
    assert 'UNMODIFIED reference' in line['cpu_baseline']['sample'] and line['cpu_baseline']['parts']['iters'] == 8
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['higher_is_better'] is True and line['unit'] == bench.UNIT
    # the GPU arm prints the identical config object for the same workload
    assert line['config'] == bench.bench_config('C4', 700, 128, 1000, 1)
    # not enough host memory for the reference's n x n matrices: labelled fallback to the port
    monkeypatch.setattr(bench, 'available_host_bytes', lambda: 1 << 20)
    r = bench.reference_measure('C4', 700, 1000, steps=1)
    assert r['kind'] == 'port' and r['sample'].startswith('FALLBACK (host memory') and r['value'] > 0
