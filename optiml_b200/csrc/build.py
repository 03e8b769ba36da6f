"""Build libsvmb200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(os.path.dirname(HERE), '_lib')
LIB = os.path.join(LIB_DIR, 'libsvmb200.so')
SOURCES = ['api.cu', 'pg.cu', 'gram.cu', 'comm.cu', 'hostmath.cu', 'devmath.cu']
HEADERS = sorted(f for f in os.listdir(HERE) if f.endswith('.cuh')) + [os.path.join('..', '..', 'include', 'svmb200.h')]  # every header
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-Wall', '--fmad=true',
              '-Xcompiler', '-ffp-contract=off']  # host arithmetic (start point, variance) rounds like NumPy


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False, defines=(), lib_dir=None):
    """``defines`` / ``lib_dir``: a variant of the library (``-DNAME=VALUE`` flags) built into its own directory --
    used by the tuning sweeps in scripts/, loaded through the SVMB200_LIB environment variable."""
    lib_dir = lib_dir or LIB_DIR
    lib = os.path.join(lib_dir, 'libsvmb200.so')
    os.makedirs(lib_dir, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    hdrs = [os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs = []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(lib_dir, src.replace('.cu', '.o'))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + [f'-D{d}' for d in defines] + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
            if verbose:
                print(' '.join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    if force or _stale(lib, objs):
        cmd = [nvcc, '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart', '-ldl', '-lpthread']
        if verbose:
            print(' '.join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return lib


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
