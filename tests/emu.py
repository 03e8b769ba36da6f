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


_cached = {}


def load(defines=()):
    """The emulated library with the prototypes of include/svmb200.h bound (built on first use).  ``defines``: a
    variant built with other shape knobs (-DSVMB200_MULTI_R=... etc.)."""
    key = tuple(defines)
    if key not in _cached:
        from optiml_b200 import _native as N
        lib = N.bind_prototypes(C.CDLL(_builder().build(defines=key)))
        lib.emu_set_schedule.argtypes = [C.c_int, C.c_uint]
        lib.emu_set_schedule.restype = None
        lib.emu_launches.restype = C.c_uint64
        lib.emu_blocks.restype = C.c_uint64
        lib.emu_allgather_calls.restype = C.c_uint64
        _cached[key] = lib
    return _cached[key]


@contextlib.contextmanager
def emulated_device(order=0, seed=1, defines=()):
    """Inside the block every call of the host mirror lands in the emulated library.  ``order``: how the emulated
    threads of a block are resumed between barriers and in which order the blocks of a launch run (0 index order,
    1 reversed, 2 seeded shuffle)."""
    from optiml_b200 import _native as N, runtime
    lib = load(defines)
    lib.emu_clear_error()
    lib.emu_set_schedule(int(order), int(seed))
    saved_lib, saved_ctx = N._lib, runtime._default_ctx
    N._lib, runtime._default_ctx = lib, None
    try:
        yield lib
    finally:
        # device objects that sit in reference cycles (estimator <-> optimizer <-> bound callback) are finalised by the cyclic
        # collector, whenever it runs: run it now, while their buffers still belong to THIS library and this context
        import gc
        gc.collect()
        ctx = runtime._default_ctx
        if ctx is not None:
            ctx.trim()
            ctx._finalizer()
        N._lib, runtime._default_ctx = saved_lib, saved_ctx
        lib.emu_set_schedule(0, 1)
