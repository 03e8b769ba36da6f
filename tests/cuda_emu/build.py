"""Build the host emulation of libsvmb200 for the CPU test-suite  --  TEST INFRASTRUCTURE ONLY.

``pg.cu`` (all solver kernels and their drivers), ``api.cu``, ``comm.cu`` and ``hostmath.cu`` are compiled FROM THE PRODUCT SOURCES
with g++ against the stand-in ``include/cuda_runtime.h``; the only source transformation is the launch syntax,
``k<<<grid, block, smem, stream>>>(args)`` -> ``emu::launch(grid, block, [=]() { k(args); })``, applied to a scratch
copy under ``_build/`` (git-ignored).  ``gram.cu`` is replaced by ``emu_standins.cpp``; NCCL and CUDA IPC by in-process
stand-ins (ranks are threads).  Nothing under
``optiml_b200/`` knows about this library; tests load it explicitly.
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'optiml_b200', 'csrc')
BUILD = os.path.join(HERE, '_build')
LIB = os.path.join(BUILD, 'libsvmb200_emu.so')
PRODUCT_SOURCES = ['pg.cu', 'api.cu', 'comm.cu', 'hostmath.cu']
HARNESS_SOURCES = ['emu_runtime.cpp', 'emu_standins.cpp']
CXXFLAGS = ['-O1', '-g', '-std=c++17', '-fPIC', '-ffp-contract=off', '-fno-omit-frame-pointer', '-DSVMB200_HOST_EMULATION',
            '-Wall', '-Wno-unknown-pragmas', '-Wno-unused-function', '-Wno-unused-variable']

_LAUNCH = re.compile(r'([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\)\s*;', re.S)


def _split_top_level(text):
    parts, depth, cur = [], 0, ''
    for ch in text:
        if ch in '([{':
            depth += 1
        elif ch in ')]}':
            depth -= 1
        if ch == ',' and depth == 0:
            parts.append(cur.strip())
            cur = ''
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(source):
    """CUDA launch syntax -> emu::launch; returns (text, number of launches rewritten)."""
    def sub(m):
        kernel, config, args = m.group(1), _split_top_level(m.group(2)), m.group(3)
        return f'emu::launch(dim3({config[0]}), dim3({config[1]}), [=]() {{ {kernel}({args}); }});'
    return _LAUNCH.subn(sub, source)


def build(force=False):
    os.makedirs(BUILD, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in PRODUCT_SOURCES + ['common.cuh', 'al_math.cuh']] + \
           [os.path.join(HERE, f) for f in HARNESS_SOURCES + ['build.py', os.path.join('include', 'cuda_runtime.h')]] + \
           [os.path.join(ROOT, 'include', 'svmb200.h')]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    units = []
    for name in PRODUCT_SOURCES:
        text, count = rewrite_launches(open(os.path.join(CSRC, name)).read())
        if name == 'pg.cu' and count < 8:
            raise RuntimeError(f'only {count} kernel launches recognised in pg.cu')
        out = os.path.join(BUILD, name.replace('.cu', '_emu.cpp'))
        with open(out, 'w') as fh:
            fh.write(f'#line 1 "{os.path.join(CSRC, name)}"\n' + text)
        units.append(out)
    units += [os.path.join(HERE, f) for f in HARNESS_SOURCES]
    cmd = ['g++'] + CXXFLAGS + ['-I', os.path.join(HERE, 'include'), '-I', CSRC, '-shared', '-o', LIB] + units + ['-lpthread']
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force=True))
