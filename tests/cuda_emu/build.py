"""Build the host emulation of libsvmb200 for the CPU test-suite  --  TEST INFRASTRUCTURE ONLY.

``pg.cu`` (all solver kernels and their drivers), ``gram.cu``, ``api.cu``, ``comm.cu`` and ``hostmath.cu`` are compiled FROM THE PRODUCT SOURCES
with g++ against the stand-in ``include/cuda_runtime.h``; the only source transformation is the launch syntax,
``k<<<grid, block, smem, stream>>>(args)`` -> ``emu::launch(grid, block, [=]() { k(args); })``, applied to a scratch
copy under ``_build/`` (git-ignored).  ``gram.cu`` (K1) is compiled too: its PTX wrappers have emulation twins (mbarriers,
tiled copies with the 128-byte swizzle, the m8n8k4 FP64 tensor-core product as a warp collective, named barriers).  NCCL
and CUDA IPC are in-process stand-ins (ranks are threads).  Nothing under
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
PRODUCT_SOURCES = ['pg.cu', 'gram.cu', 'api.cu', 'comm.cu', 'hostmath.cu', 'devmath.cu']
HARNESS_SOURCES = ['emu_runtime.cpp']
# -fno-gnu-unique / -Bsymbolic: several variants of the library can live in one process (shape sweeps); the statics of
# template kernels ("__shared__" arrays whose size depends on the shape) must not be merged across them
CXXFLAGS = ['-O1', '-g', '-std=c++17', '-fPIC', '-ffp-contract=off', '-fno-omit-frame-pointer', '-DSVMB200_HOST_EMULATION',
            '-fno-gnu-unique',
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
        smem = config[2] if len(config) > 2 else '0'
        return f'emu::launch(dim3({config[0]}), dim3({config[1]}), (size_t)({smem}), [=]() {{ {kernel}({args}); }});'
    text, count = _LAUNCH.subn(sub, source)
    return text.replace('__noinline__', 'EMU_NOINLINE'), count


def _stale(target, deps):
    return not os.path.exists(target) or os.path.getmtime(target) < max(os.path.getmtime(d) for d in deps)


def build(force=False, defines=()):
    """The emulated library; ``defines``: a variant (the shape knobs of the multi-vector pass, which only pg.cu reads)
    under its own name -- the other objects are compiled once and shared."""
    os.makedirs(BUILD, exist_ok=True)
    defines = tuple(defines)
    suffix = ''.join('_' + d.replace('SVMB200_', '').replace('=', '') for d in defines)
    lib = os.path.join(BUILD, f'libsvmb200_emu{suffix}.so')
    common = [os.path.join(CSRC, h) for h in sorted(os.listdir(CSRC)) if h.endswith('.cuh')] + \
             [os.path.join(ROOT, 'include', 'svmb200.h'),
              os.path.join(HERE, 'include', 'cuda_runtime.h'), os.path.join(HERE, 'include', 'cudaTypedefs.h'),
              os.path.abspath(__file__)]
    include = ['-I', os.path.join(HERE, 'include'), '-I', CSRC]
    jobs, objects = [], []
    for name in PRODUCT_SOURCES:
        src = os.path.join(CSRC, name)
        variant = suffix if name == 'pg.cu' else ''
        obj = os.path.join(BUILD, name.replace('.cu', f'_emu{variant}.o'))
        objects.append(obj)
        if force or _stale(obj, [src] + common):
            text, count = rewrite_launches(open(src).read())
            if name == 'pg.cu' and count + text.count('svm_launch_chained(') < 8:
                raise RuntimeError(f'only {count} kernel launches recognised in pg.cu')
            cpp = obj[:-2] + '.cpp'
            with open(cpp, 'w') as fh:
                fh.write(f'#line 1 "{src}"\n' + text)
            flags = [f'-D{d}' for d in defines] if name == 'pg.cu' else []
            jobs.append(['g++'] + CXXFLAGS + flags + include + ['-c', cpp, '-o', obj])
    for name in HARNESS_SOURCES:
        src = os.path.join(HERE, name)
        obj = os.path.join(BUILD, name.replace('.cpp', '.o'))
        objects.append(obj)
        if force or _stale(obj, [src] + common):
            jobs.append(['g++'] + CXXFLAGS + include + ['-c', src, '-o', obj])
    procs = [subprocess.Popen(cmd) for cmd in jobs]   # a handful of translation units: compile them side by side
    if any(p.wait() != 0 for p in procs):
        raise RuntimeError('tests/cuda_emu: compilation failed')
    if force or jobs or _stale(lib, objects):
        subprocess.run(['g++', '-shared', '-fno-gnu-unique', '-Wl,-Bsymbolic', '-o', lib] + objects + ['-lpthread'], check=True)
    return lib


if __name__ == '__main__':
    print(build(force=True))
