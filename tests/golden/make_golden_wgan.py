#!/usr/bin/env python
"""Generates tests/golden/wgan_gp.npz by running the reference script's own Generator / Critic classes
(conditional_gan/mnist/mnist_wgan_conditional.py:51-108, AST-lifted from /root/reference: the script trains at import time)
and its loop body (:132-168, executed statement by statement as written in tests/test_wgan_oracle_vs_reference.py, the
draws of :139 / :144 / :160-161 injected) on CPU, at reduced widths (critic 64, generator 64, hidden 32, batch 8,
n_critic 2, 3 iterations).  Run from the repository root in the build container:

    python tests/golden/make_golden_wgan.py

Inputs and initial parameters are NOT stored (``oracle.wgan_gp.synth_params`` / ``synth_batch`` regenerate them
bit-exactly from integer seeds).  Stored: the losses of every iteration and, for every tensor of both state_dicts after
the run, [sum, sum of absolute values, 32 fixed samples]."""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import wgan_gp as O  # noqa: E402
from tests.test_wgan_oracle_vs_reference import reference_run  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CFG = dict(batchsize=8, latent_size=16, n_critic=2, critic_size=64, generator_size=64, critic_hidden_size=32)
STEPS, SEED0, SEED_G, SEED_C = 3, 31, 7, 8


def summary(t):
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, 32).long()
    return np.concatenate([[t.sum().item(), t.abs().sum().item()], t[idx].numpy()])


def main():
    hp = O.Hyper(**CFG)
    params = (O.synth_params(O.g_shapes(hp), SEED_G), O.synth_params(O.c_shapes(hp), SEED_C))
    critic, generator, log, _ = reference_run(hp, STEPS, SEED0, params)
    losses = np.array([[r["critic_loss"], r["gp"], r.get("generator_loss", np.nan)] for r in log], dtype=np.float64)
    out = OrderedDict(meta=np.array([STEPS, SEED0, SEED_G, SEED_C] + [CFG[k] for k in sorted(CFG)]), losses=losses)
    for pre, net in (("G.", generator), ("C.", critic)):
        for k, v in net.state_dict().items():
            out[pre + k] = summary(v)
    np.savez_compressed(os.path.join(OUT, "wgan_gp.npz"), **out)
    print("wgan_gp", losses.tolist(), len(out) - 2, "tensors")


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
