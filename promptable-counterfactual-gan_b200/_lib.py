"""ctypes binding of libpcg.so (the C ABI declared in include/pcg.h).

There is no CPU fallback: if the shared library is missing or a CUDA device is absent, every
compute entry point raises.  ``load()`` is cheap to call repeatedly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpcg.so")
_lib = None


class PcgError(RuntimeError):
    pass


def build(verbose=False):
    """Compiles csrc/*.cu for sm_100a into csrc/libpcg.so (nvcc cross-compiles without a GPU)."""
    import subprocess
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise PcgError("building libpcg.so failed")


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PcgError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.pcg_last_error.restype = ctypes.c_char_p
        _lib.pcg_launch_count.restype = ctypes.c_ulonglong
    return _lib


def check(rc):
    if rc != 0:
        raise PcgError(f"libpcg error {rc}: {load().pcg_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (or NULL for None) as c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(load().pcg_launch_count())
