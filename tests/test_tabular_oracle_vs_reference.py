"""Pins oracle/tabular_countergan.py against the reference's own, unmodified ``train_countergan`` functions
(moons and KC house sales; build container only).  The trainers are imported from /root/reference and run on CPU;
their data loader is replaced by a list of fixed batches, ``torch.randint`` (targets, masks) and
``Tensor.exponential_`` (the noise inside torch's own ``F.gumbel_softmax``) return injected values, and the
Discriminator the trainer constructs internally is given known initial weights / power-iteration vectors."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import tabular_countergan as T
from tests._refload import experiment, have_reference

pytestmark = pytest.mark.reference


class _Loader:
    def __init__(self, batches):
        self.b = batches

    def __iter__(self):
        return iter(self.b)

    def __len__(self):
        return len(self.b)


def _bn_shadowed(k):
    """Linear biases feeding a train-mode BatchNorm: zero true gradient, Adam turns rounding noise into a +-lr walk
    (see oracle.mnist_countergan.is_bn_shadowed_bias)."""
    return k in ("net.0.bias", "net.3.bias", "net.6.bias") or k.endswith("fc1.bias") or k.endswith("fc2.bias")


def _close_params(ref_sd, mine, lr, steps, tag):
    for k, v in ref_sd.items():
        if "num_batches" in k:
            continue
        if tag == "G" and _bn_shadowed(k):
            assert (v.float() - mine[k].detach().float()).abs().max() <= 2.02 * lr * steps, (tag, k)
            continue
        m = mine[k].detach().float()
        v = v.float()
        if "running" in k or k.endswith("_u") or k.endswith("_v"):
            # running means follow the +-lr walk of the BN-shadowed biases in front of them
            assert torch.allclose(v, m, atol=2.02 * lr * steps, rtol=2e-3), (tag, k, (v - m).abs().max())
        else:
            d = (v - m).abs()
            assert d.max() <= 2.02 * lr * steps and d.mean() <= 0.02 * lr, (tag, k, d.max().item(), d.mean().item())


@pytest.mark.skipif(not have_reference(), reason="no reference")
def test_moons_train_countergan_matches_oracle(tmp_path, monkeypatch):
    with experiment("conditional_counteRGAN/moons") as imp:
        trainer = imp("trainer")
        gen, dis, clf = imp("models.generator"), imp("models.discriminator"), imp("models.nn_classifier")
    gs, ds, cs = T.moons_shapes()
    PG, PD, PC = T.synth_params(gs, 1), T.synth_params(ds, 2), T.synth_params(cs, 3)
    BD = T.sn_buffers(T.moons_d_dims(), 4)
    G = gen.ResidualGenerator(2, 32, 3)
    G.load_state_dict({**PG, **T.bn_buffers(gs)})
    C = clf.NNClassifier(2)
    C.load_state_dict(PC)
    made = {}

    def make_d(*a, **k):
        D = dis.Discriminator(*a, **k)
        D.load_state_dict({**PD, **BD})
        made["D"] = D
        return D
    B, steps = 16, 3
    batches = [T.moons_batch(B, 20 + i) for i in range(steps)]
    draws = []
    for b in batches:
        draws += [b[2], b[3].long()]
    it = iter(draws)
    monkeypatch.setattr(trainer, "Discriminator", make_d)
    monkeypatch.setattr(trainer, "DataLoader", lambda *a, **k: _Loader([(b[0], b[1]) for b in batches]))
    monkeypatch.setattr(torch, "randint", lambda *a, **k: next(it).clone())
    cfg = {"cuda": "cpu", "seed": 42, "epochs": 1, "batch_size": B, "lr_G": 1e-3, "lr_D": 1e-3, "lambda_cls": 2.0,
           "lambda_reg_l1": 5.0, "lambda_reg_l2": 5.0, "lambda_mask": 3.0, "input_dim": 2, "hidden_dim": 32,
           "out_dir": str(tmp_path), "generator_path": str(tmp_path / "g.pt")}
    X = torch.cat([b[0] for b in batches]).numpy()
    y = torch.cat([b[1] for b in batches]).numpy()
    y[:3] = [0, 1, 2]       # np.unique(y_train).size must be 3
    trainer.train_countergan(G, cfg, X, y, C)
    monkeypatch.undo()
    S = T.make_state(PG, T.bn_buffers(gs), PD, BD, PC)
    for b in batches:
        sc, _ = T.moons_step(S, *b)
    assert all(np.isfinite(list(sc.values())))
    _close_params(torch.load(cfg["generator_path"]), {**S["G"], **S["GB"]}, 1e-3, steps, "G")
    _close_params(made["D"].state_dict(), {**S["D"], **S["DB"]}, 1e-3, steps, "D")


@pytest.mark.skipif(not have_reference(), reason="no reference")
def test_kc_train_countergan_matches_oracle(tmp_path, monkeypatch):
    cwd = os.getcwd()
    os.chdir(tmp_path)          # the reference config.py creates results/ in the CWD at import time
    try:
        with experiment("conditional_counteRGAN/house_sales_kc_usa") as imp:
            trainer = imp("trainer")
            gen, dis, clf = imp("models.generator"), imp("models.discriminator"), imp("models.nn_classifier")
            rcfg = imp("config").config
    finally:
        os.chdir(cwd)
    assert list(rcfg["categorical_info"].keys()) == list(T.KC_CAT.keys())
    assert rcfg["continuous_idx"] == T.KC_CONT and sorted(rcfg["immutable_idx"]) == sorted(T.KC_IMMUTABLE)
    gs, ds, cs = T.kc_shapes()
    PG, PD, PC = T.synth_params(gs, 1), T.synth_params(ds, 2), T.synth_params(cs, 3)
    BD, BC = T.sn_buffers(T.kc_d_dims(), 4), T.bn_buffers(cs, 5, randomize=True)
    cat_info = {k: {"n": v["n"], "raw_values": v["raw_values"]} for k, v in rcfg["categorical_info"].items()}
    G = gen.ResidualGenerator(17, 32, 4, continuous_idx=rcfg["continuous_idx"], categorical_info=cat_info, tau=0.5)
    assert [k for k, _ in G.named_parameters()] == list(gs.keys())
    G.load_state_dict({**PG, **T.bn_buffers(gs)})
    C = clf.NNClassifier(17, output_dim=4)
    assert [k for k, _ in C.named_parameters()] == list(cs.keys())
    C.load_state_dict({**PC, **BC})
    made = {}

    def make_d(*a, **k):
        D = dis.Discriminator(*a, **k)
        assert [n for n, _ in D.named_parameters()] == list(ds.keys())
        D.load_state_dict({**PD, **BD})
        made["D"] = D
        return D
    B, steps = 16, 2
    batches = [T.kc_batch(B, 30 + i) for i in range(steps)]
    draws, noise = [], []
    for b in batches:
        draws += [b[2], b[3].long()]
        noise += b[4]
    it, ni = iter(draws), iter(noise)
    monkeypatch.setattr(trainer, "Discriminator", make_d)
    monkeypatch.setattr(trainer, "DataLoader", lambda *a, **k: _Loader([(b[0], b[1]) for b in batches]))
    monkeypatch.setattr(torch, "randint", lambda *a, **k: next(it).clone())
    monkeypatch.setattr(torch.Tensor, "exponential_", lambda self, *a, **k: self.copy_(next(ni)))
    cfg = dict(rcfg, cuda="cpu", epochs=1, batch_size=B, scaler=None, out_dir=str(tmp_path),
               generator_path=str(tmp_path / "g.pt"))
    X = torch.cat([b[0] for b in batches]).numpy()
    y = torch.cat([b[1] for b in batches]).numpy()
    y[:4] = [0, 1, 2, 3]
    trainer.train_countergan(G, cfg, X, y, C.eval())
    monkeypatch.undo()
    S = T.make_state(PG, T.bn_buffers(gs), PD, BD, PC, BC)
    nv = T.kc_norm_vals()
    for b in batches:
        sc, _ = T.kc_step(S, *b, nv)
    assert all(np.isfinite(list(sc.values())))
    _close_params(torch.load(cfg["generator_path"]), {**S["G"], **S["GB"]}, 1e-3, steps, "G")
    _close_params(made["D"].state_dict(), {**S["D"], **S["DB"]}, 1e-3, steps, "D")
