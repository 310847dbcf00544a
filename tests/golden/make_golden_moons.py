#!/usr/bin/env python
"""Generates tests/golden/moons_{cgan,gan}.npz by running the UNMODIFIED reference code of the two moons GANs
(AST-lifted from /root/reference: the scripts train at import time and cannot be imported) on seeded inputs with
the random draws (torch.randn / torch.randint / numpy shuffles) replaced by injected values.  Run from the repository
root in the build container:

    python tests/golden/make_golden_moons.py

Inputs are NOT stored: batches come bit-exactly from ``oracle.moons_gan.synth_batch`` seeds; the initial parameters are
stored (they are the reference's own ``nn.Linear`` initialisation under ``torch.manual_seed``).  Stored outputs: the
parameters of both nets after the run and the per-epoch loss sums the reference accumulated."""
import os
import sys
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import moons_gan as M  # noqa: E402
from tests._refload import lift  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def pack(prefix, sd):
    return {prefix + k.replace("net.", ""): v.detach().numpy().copy() for k, v in sd.items()}


def cgan(B=48, nb=3, seed0=60):
    ns, loop = lift("conditional_gan/moons/make_moons_cgan.py", ("Generator", "Discriminator", "one_hot_encode"),
                    loop_var="epoch")
    cfg = {"n_samples": B * nb, "z_dim": 32, "hidden_dim": 128, "label_dim": 2, "batch_size": B, "lr": 1e-3, "epochs": 1}
    torch.manual_seed(1)
    G, D = ns["Generator"](32, 2, 128), ns["Discriminator"](2, 128)
    out = {"meta": np.array([B, nb, seed0, 2, 1])}
    out.update(pack("G0.", G.state_dict()))
    out.update(pack("D0.", D.state_dict()))
    batches = [M.synth_batch(B, seed0 + i) for i in range(nb)]
    ns.update(config=cfg, generator=G, discriminator=D,
              optimizer_G=torch.optim.Adam(G.parameters(), lr=1e-3), optimizer_D=torch.optim.Adam(D.parameters(), lr=1e-3),
              real_samples=torch.cat([b[0] for b in batches]), real_labels=torch.cat([b[1].argmax(1) for b in batches]),
              loss_D_values=[], loss_G_values=[])
    zs, labs = [], []
    for b in batches:
        zs += [b[2], b[4]]
        labs += [b[3].argmax(1), b[5].argmax(1)]
    zi, li = iter(zs), iter(labs)
    with mock.patch.object(np.random, "permutation", lambda n: np.arange(n)), \
            mock.patch.object(torch, "randn", lambda *a, **k: next(zi).clone()), \
            mock.patch.object(torch, "randint", lambda *a, **k: next(li).clone()):
        exec(loop, ns)
    out.update(pack("G1.", G.state_dict()))
    out.update(pack("D1.", D.state_dict()))
    out["loss_D"] = np.array(ns["loss_D_values"], dtype=np.float64)
    out["loss_G"] = np.array(ns["loss_G_values"], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "moons_cgan.npz"), **out)
    print("moons_cgan", out["loss_D"], out["loss_G"])


def gan(B=16, nb=3, epochs=2, seed0=40):
    ns, _ = lift("simple_gan/moons/make_moons_gan.py", ("build_generator", "build_discriminator", "train_gan"))
    cfg = {"z_dim": 32, "hidden_dim": 128, "batch_size": B, "lr": 1e-3, "epochs": epochs}
    torch.manual_seed(0)
    G, D = ns["build_generator"](32, 128), ns["build_discriminator"](128)
    out = {"meta": np.array([B, nb, seed0, 0, epochs])}
    out.update(pack("G0.", G.state_dict()))
    out.update(pack("D0.", D.state_dict()))
    batches = [M.synth_batch(B, seed0 + i, label_dim=0) for i in range(epochs * nb)]
    X = np.concatenate([b[0].numpy() for b in batches[:nb]]).astype(np.float32)
    zs = []
    for b in batches:
        zs += [b[2], b[4]]
    it = iter(zs)
    with mock.patch.object(np.random, "shuffle", lambda x: None), \
            mock.patch.object(torch, "randn", lambda *a, **k: next(it).clone()):
        lossD, lossG = ns["train_gan"](X.copy(), G, D, cfg)
    out.update(pack("G1.", G.state_dict()))
    out.update(pack("D1.", D.state_dict()))
    out["loss_D"] = np.array(lossD, dtype=np.float64)
    out["loss_G"] = np.array(lossG, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "moons_gan.npz"), **out)
    print("moons_gan", out["loss_D"], out["loss_G"])


if __name__ == "__main__":
    torch.set_num_threads(4)
    cgan()
    gan()
