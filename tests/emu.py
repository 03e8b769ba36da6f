"""Run the host mirror against the HOST EMULATION of libsvmb200 (tests/cuda_emu)  --  TEST INFRASTRUCTURE ONLY.

``emulated_device()`` swaps the loaded library of ``optiml_b200._native`` for the emulated one for the duration of a
test (and restores it afterwards); the product code is untouched and has no way to select the emulation itself.
"""
import contextlib
import ctypes as C
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def _builder():
    spec = importlib.util.spec_from_file_location('cuda_emu_build', os.path.join(HERE, 'cuda_emu', 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_cached = None


def load():
    """The emulated library with the prototypes of include/svmb200.h bound (built on first use)."""
    global _cached
    if _cached is None:
        from optiml_b200 import _native as N
        lib = N.bind_prototypes(C.CDLL(_builder().build()))
        lib.emu_set_schedule.argtypes = [C.c_int, C.c_uint]
        lib.emu_set_schedule.restype = None
        lib.emu_launches.restype = C.c_uint64
        lib.emu_blocks.restype = C.c_uint64
        lib.emu_allgather_calls.restype = C.c_uint64
        _cached = lib
    return _cached


@contextlib.contextmanager
def emulated_device(order=0, seed=1):
    """Inside the block every call of the host mirror lands in the emulated library.  ``order``: how the emulated
    threads of a block are resumed between barriers (0 index order, 1 alternating reversed, 2 seeded shuffle)."""
    from optiml_b200 import _native as N, runtime
    lib = load()
    lib.emu_clear_error()
    lib.emu_set_schedule(int(order), int(seed))
    saved_lib, saved_ctx = N._lib, runtime._default_ctx
    N._lib, runtime._default_ctx = lib, None
    try:
        yield lib
    finally:
        ctx = runtime._default_ctx
        if ctx is not None:
            ctx.trim()
            ctx._finalizer()
        N._lib, runtime._default_ctx = saved_lib, saved_ctx
        lib.emu_set_schedule(0, 1)
