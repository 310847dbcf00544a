#!/usr/bin/env python
"""Device time of pcg_linear_wgrad_small against pcg_conv_wgrad + pcg_colsum on the tabular layer shapes (a CUDA graph of 20
launches each, so host latency is out of the number).   python tools/bench_wgrad_small.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcg_b200  # noqa: E402,F401
from pcg_b200 import graphs, ops as K  # noqa: E402


def timed(fn, reps=20):
    fn()
    g = graphs.capture(lambda: [fn() for _ in range(reps)])
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


for M, Kd, N in [(4096, 21, 32), (4096, 38, 32), (4096, 32, 32), (4096, 32, 64), (4096, 64, 128), (4096, 128, 1), (4096, 32, 10),
                 (64, 7, 32), (64, 32, 32), (64, 32, 16), (64, 16, 2), (128, 17, 256)]:
    x, dy = torch.randn(M, Kd, device="cuda"), torch.randn(M, N, device="cuda")
    dw, db = torch.empty(N, Kd, device="cuda"), torch.empty(N, device="cuda")
    need = K.linear_wgrad_small_scratch_floats(M, Kd, N)
    line = f"M={M:5d} K={Kd:3d} N={N:3d}: "
    if need > 0:
        s1 = torch.zeros(need, device="cuda")
        line += f"one call {timed(lambda: K.linear_wgrad_small(x, dy, s1, dw, db)):6.1f} us   "
    s2, st = K.conv_wgrad_scratch(M, 1, 1, Kd, N, 1, 1, 0, "cuda"), K.stat_scratch(max(N, 4), "cuda")

    def old():
        K.conv_wgrad(x, dy, M, 1, 1, Kd, N, 1, 1, 0, s2, dw)
        K.colsum(dy, st, db)
    print(line + f"conv_wgrad + colsum {timed(old):6.1f} us", flush=True)
