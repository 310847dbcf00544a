#!/usr/bin/env python
"""Times the conditional WGAN-GP iteration (pcg_b200.wgan) at the reference's sizes: batch 128, widths 1024, the loop's
schedule of n_critic = 5 batches (5 critic updates + 1 generator update), as CUDA graphs.  Prints ms per batch and
samples/s for the exact fp32 mode and the tensor-core modes, plus the per-launcher breakdown of one eager cycle."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcg_b200  # noqa: E402,F401
from pcg_b200 import _lib  # noqa: E402
from pcg_b200.wgan import Hyperparameter, WganGpPlan  # noqa: E402
from oracle import wgan_gp as O  # noqa: E402


def main():
    hp = Hyperparameter()
    ohp = O.Hyper()
    B = hp.batchsize
    PG, PC = O.synth_params(O.g_shapes(ohp), 7), O.synth_params(O.c_shapes(ohp), 8)
    b = {k: v.cuda() for k, v in O.synth_batch(ohp, B, 3).items()}
    for name, tc, terms in (("fp32", False, 1), ("bf16x3", True, 3), ("bf16", True, 1)):
        plan = WganGpPlan(hp, B, "cuda", tensor_cores=tc, operand_terms=terms)
        plan.G.load(PG)
        plan.C.load(PC)
        plan.refresh()
        plan.load_inputs(b["real"], b["labels"], b["noise"], b["alpha"], b["labels_g"], b["noise_g"])
        for w in (True, False):
            plan.run(w)
        torch.cuda.synchronize()
        cycles = 4
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(cycles):
            for i in range(hp.n_critic):
                plan.run(i == 0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (cycles * hp.n_critic)
        ec, eg = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ec.record()
        for _ in range(5):
            plan.run(False)
        eg.record()
        torch.cuda.synchronize()
        print(f"{name}: {ms:.3f} ms/batch  {B / ms * 1e3:.0f} samples/s  critic-only {ec.elapsed_time(eg) / 5:.3f} ms  "
              f"scal {[round(v, 4) for v in plan.scal.tolist()[:5]]}", flush=True)
        if os.environ.get("PCG_WGAN_BREAKDOWN", "1") == "1":
            plan.use_graph = False
            L = _lib.load()
            L.pcg_profile_begin()
            plan.run(True)
            import ctypes
            import json
            buf = ctypes.create_string_buffer(1 << 16)
            _lib.check(L.pcg_profile_end(buf, ctypes.c_size_t(len(buf))))
            prof = json.loads(buf.value.decode())
            print("   " + "  ".join(f"{k}={v['ms']:.3f}ms/{v['launches']}"
                                    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:14]), flush=True)
        del plan
        torch.cuda.empty_cache()


def torch_eager():
    """The same iteration through torch autograd on this GPU (the oracle's functional restatement moved to cuda: F.conv2d /
    conv_transpose2d, autograd.grad(create_graph=True), default TF32-off fp32 math), batch 128, n_critic schedule."""
    from collections import OrderedDict
    ohp = O.Hyper()
    B = ohp.batchsize
    PG, PC = O.synth_params(O.g_shapes(ohp), 7), O.synth_params(O.c_shapes(ohp), 8)
    S = O.make_state(PG, O.g_buffers(ohp), PC)
    for k in ("G", "GB", "C"):
        S[k] = OrderedDict((n, t.detach().cuda().requires_grad_(t.requires_grad)) for n, t in S[k].items())
    S["adam_g"], S["adam_c"] = O.adam_init(S["G"]), O.adam_init(S["C"])
    b = {k: v.cuda() for k, v in O.synth_batch(ohp, B, 3).items()}
    eye = torch.eye(ohp.num_classes, device="cuda")

    def step(i):
        O.critic_step(S, ohp, b["real"], eye[b["labels"]], b["noise"], b["alpha"])
        if i % ohp.n_critic == 0:
            O.generator_step(S, ohp, b["noise_g"], eye[b["labels_g"]])
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        for i in range(5):
            step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 20
        for i in range(n):
            step(i)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / n * 1e3
        print(f"torch eager (autograd, tf32={tf32}): {ms:.3f} ms/batch  {B / ms * 1e3:.0f} samples/s", flush=True)


if __name__ == "__main__":
    main()
    if os.environ.get("PCG_WGAN_TORCH", "1") == "1":
        torch_eager()
