"""Pins oracle/moons_gan.py against the reference scripts' own code (build container only): class / function
definitions and the training-loop statement are AST-lifted from the reference files and executed unmodified on
CPU with the random draws (torch.randn / torch.randint / numpy shuffles) replaced by injected values."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import moons_gan as M
from tests._refload import have_reference, lift

pytestmark = pytest.mark.reference


def _strip(sd):
    return OrderedDict((k.replace("net.", ""), v) for k, v in sd.items())


@pytest.mark.skipif(not have_reference(), reason="no reference")
def test_simple_gan_train_gan_matches_oracle(monkeypatch):
    ns, _ = lift("simple_gan/moons/make_moons_gan.py", ("build_generator", "build_discriminator", "train_gan"))
    cfg = {"z_dim": 32, "hidden_dim": 128, "batch_size": 16, "lr": 1e-3, "epochs": 2}
    torch.manual_seed(0)
    G, D = ns["build_generator"](32, 128), ns["build_discriminator"](128)
    S = M.make_state(_strip(G.state_dict()), _strip(D.state_dict()))
    n_batches = 3
    batches = [M.synth_batch(16, 40 + i, label_dim=0) for i in range(cfg["epochs"] * n_batches)]
    X = np.concatenate([b[0].numpy() for b in batches[:n_batches]]).astype(np.float32)
    # epoch e sees the same X (shuffle disabled) -> batches of epoch 1 reuse the reals of epoch 0
    for e in range(1, cfg["epochs"]):
        for i in range(n_batches):
            b = list(batches[e * n_batches + i])
            b[0] = batches[i][0]
            batches[e * n_batches + i] = tuple(b)
    zs = []
    for b in batches:
        zs += [b[2], b[4]]
    it = iter(zs)
    monkeypatch.setattr(np.random, "shuffle", lambda x: None)
    monkeypatch.setattr(torch, "randn", lambda *a, **k: next(it).clone())
    lossD, lossG = ns["train_gan"](X.copy(), G, D, cfg)
    monkeypatch.undo()
    od, og = [], []
    for e in range(cfg["epochs"]):
        d = g = 0.0
        for i in range(n_batches):
            ld, lg, _ = M.gan_step(S, *batches[e * n_batches + i])
            d, g = d + ld, g + lg
        od.append(d)
        og.append(g)
    assert np.allclose(lossD, od, rtol=1e-5) and np.allclose(lossG, og, rtol=1e-5)
    for k, v in _strip(G.state_dict()).items():
        assert torch.allclose(v, S["G"][k], atol=2e-6, rtol=1e-4), k
    for k, v in _strip(D.state_dict()).items():
        assert torch.allclose(v, S["D"][k], atol=2e-6, rtol=1e-4), k


@pytest.mark.skipif(not have_reference(), reason="no reference")
def test_cgan_script_loop_matches_oracle(monkeypatch):
    ns, loop = lift("conditional_gan/moons/make_moons_cgan.py", ("Generator", "Discriminator", "one_hot_encode"),
                    loop_var="epoch")
    assert loop is not None
    B, nb = 16, 3
    cfg = {"n_samples": B * nb, "z_dim": 32, "hidden_dim": 128, "label_dim": 2, "batch_size": B, "lr": 1e-3, "epochs": 1}
    torch.manual_seed(1)
    G, D = ns["Generator"](32, 2, 128), ns["Discriminator"](2, 128)
    S = M.make_state(_strip(G.state_dict()), _strip(D.state_dict()))
    batches = [M.synth_batch(B, 60 + i) for i in range(nb)]
    ns.update(config=cfg, generator=G, discriminator=D,
              optimizer_G=torch.optim.Adam(G.parameters(), lr=1e-3), optimizer_D=torch.optim.Adam(D.parameters(), lr=1e-3),
              real_samples=torch.cat([b[0] for b in batches]), real_labels=torch.cat([b[1].argmax(1) for b in batches]),
              loss_D_values=[], loss_G_values=[])
    zs, labs = [], []
    for b in batches:
        zs += [b[2], b[4]]
        labs += [b[3].argmax(1), b[5].argmax(1)]
    zi, li = iter(zs), iter(labs)
    monkeypatch.setattr(np.random, "permutation", lambda n: np.arange(n))
    monkeypatch.setattr(torch, "randn", lambda *a, **k: next(zi).clone())
    monkeypatch.setattr(torch, "randint", lambda *a, **k: next(li).clone())
    exec(loop, ns)
    monkeypatch.undo()
    d = g = 0.0
    for b in batches:
        ld, lg, _ = M.gan_step(S, *b)
        d, g = d + ld, g + lg
    assert abs(ns["loss_D_values"][0] - d) < 1e-4 * abs(d) and abs(ns["loss_G_values"][0] - g) < 1e-4 * abs(g)
    for k, v in _strip(G.state_dict()).items():
        assert torch.allclose(v, S["G"][k], atol=2e-6, rtol=1e-4), k
    for k, v in _strip(D.state_dict()).items():
        assert torch.allclose(v, S["D"][k], atol=2e-6, rtol=1e-4), k
