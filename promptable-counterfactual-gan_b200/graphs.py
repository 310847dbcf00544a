"""CUDA-graph capture for the step plans, in one place.

Why a helper: while a stream of this thread is capturing (``capture_error_mode="thread_local"``), CUDA forbids the
"unsafe" runtime calls — cudaFree, cudaGraphExecDestroy, cudaMalloc, synchronisation — and one such call invalidates the
capture (``cudaErrorStreamCaptureInvalidated`` at ``capture_end``).  Python's *cyclic* garbage collector can run at any
allocation, and the garbage it finds may be an older plan (plan <-> GraphStep reference cycle) whose ``CUDAGraph`` /
native plan destructor makes exactly those calls.  torch.cuda.graph no longer runs ``gc.collect()`` on entry by
default, so a capture that follows other plans in the same process (the full GPU test suite, a trainer that builds a
second plan for a tail batch) was killed by a collection that happened to trigger inside it.

``capture(body)`` therefore collects *before* the capture and keeps the cyclic collector off *during* it; native
handles whose finaliser runs while a capture is in flight are parked by ``defer_destroy`` and released afterwards.
"""
import gc

import torch

_deferred = []
capture_log = []          # libpcg kernel launches recorded by each capture() so far (bench.py: launches per replay)


def capturing():
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def defer_destroy(fn):
    """Runs ``fn`` now, or after the current capture if one is in flight (finalisers must not cudaFree in a capture)."""
    if capturing():
        _deferred.append(fn)
    else:
        fn()


def flush_deferred():
    while _deferred and not capturing():
        fn = _deferred.pop()
        try:
            fn()
        except Exception:       # a finaliser must never raise into unrelated code
            pass


def capture(body, sync=True):
    """Captures ``body()`` (kernel launches on the current stream) into a new ``torch.cuda.CUDAGraph``."""
    gc.collect()
    flush_deferred()
    if sync:
        torch.cuda.synchronize()
    from . import _lib
    g = torch.cuda.CUDAGraph()
    was_enabled = gc.isenabled()
    gc.disable()
    n0 = _lib.launch_count()
    try:
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            body()
        capture_log.append(_lib.launch_count() - n0)
    finally:
        if was_enabled:
            gc.enable()
        flush_deferred()
    return g
