#!/usr/bin/env python
"""Generates tests/golden/tabular_{moons,kc}.npz by running the reference's own, UNMODIFIED ``train_countergan``
functions (conditional_counteRGAN/moons/trainer.py:31, house_sales_kc_usa/trainer.py:186; imported from /root/reference,
matplotlib stubbed) on CPU with fixed batches and injected random draws (targets / masks via ``torch.randint``, the
Gumbel noise via ``Tensor.exponential_``).  Run from the repository root in the build container:

    python tests/golden/make_golden_tabular.py

Inputs and initial parameters are NOT stored: they come bit-exactly from integer seeds through
``oracle.tabular_countergan`` (synth_params / sn_buffers / bn_buffers / moons_batch / kc_batch).  Stored: every tensor of
the generator's and the critic's ``state_dict`` after the run."""
import os
import sys
import tempfile
from unittest import mock

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tabular_countergan as T  # noqa: E402
from tests._refload import experiment  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
MOONS = dict(B=16, steps=3, seed0=20)
KC = dict(B=16, steps=2, seed0=30)


class _Loader:
    def __init__(self, batches):
        self.b = batches

    def __iter__(self):
        return iter(self.b)

    def __len__(self):
        return len(self.b)


def _save(name, meta, g_sd, d_sd):
    out = {"meta": np.array(meta)}
    for pre, sd in (("G.", g_sd), ("D.", d_sd)):
        for k, v in sd.items():
            out[pre + k] = v.detach().cpu().numpy().copy()
    np.savez_compressed(os.path.join(OUT, f"tabular_{name}.npz"), **out)
    print(name, len(out) - 1, "tensors")


def moons(tmp):
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
    B, steps, seed0 = MOONS["B"], MOONS["steps"], MOONS["seed0"]
    batches = [T.moons_batch(B, seed0 + i) for i in range(steps)]
    draws = []
    for b in batches:
        draws += [b[2], b[3].long()]
    it = iter(draws)
    cfg = {"cuda": "cpu", "seed": 42, "epochs": 1, "batch_size": B, "lr_G": 1e-3, "lr_D": 1e-3, "lambda_cls": 2.0,
           "lambda_reg_l1": 5.0, "lambda_reg_l2": 5.0, "lambda_mask": 3.0, "input_dim": 2, "hidden_dim": 32,
           "out_dir": tmp, "generator_path": os.path.join(tmp, "g_moons.pt")}
    X = torch.cat([b[0] for b in batches]).numpy()
    y = torch.cat([b[1] for b in batches]).numpy()
    y[:3] = [0, 1, 2]       # np.unique(y_train).size must be 3 (only the count is used)
    with mock.patch.object(trainer, "Discriminator", make_d), \
            mock.patch.object(trainer, "DataLoader", lambda *a, **k: _Loader([(b[0], b[1]) for b in batches])), \
            mock.patch.object(torch, "randint", lambda *a, **k: next(it).clone()):
        trainer.train_countergan(G, cfg, X, y, C)
    _save("moons", [B, steps, seed0], torch.load(cfg["generator_path"]), made["D"].state_dict())


def kc(tmp):
    cwd = os.getcwd()
    os.chdir(tmp)           # the reference config.py creates results/ in the CWD at import time
    try:
        with experiment("conditional_counteRGAN/house_sales_kc_usa") as imp:
            trainer = imp("trainer")
            gen, dis, clf = imp("models.generator"), imp("models.discriminator"), imp("models.nn_classifier")
            rcfg = imp("config").config
    finally:
        os.chdir(cwd)
    gs, ds, cs = T.kc_shapes()
    PG, PD, PC = T.synth_params(gs, 1), T.synth_params(ds, 2), T.synth_params(cs, 3)
    BD, BC = T.sn_buffers(T.kc_d_dims(), 4), T.bn_buffers(cs, 5, randomize=True)
    cat_info = {k: {"n": v["n"], "raw_values": v["raw_values"]} for k, v in rcfg["categorical_info"].items()}
    G = gen.ResidualGenerator(17, 32, 4, continuous_idx=rcfg["continuous_idx"], categorical_info=cat_info, tau=0.5)
    G.load_state_dict({**PG, **T.bn_buffers(gs)})
    C = clf.NNClassifier(17, output_dim=4)
    C.load_state_dict({**PC, **BC})
    made = {}

    def make_d(*a, **k):
        D = dis.Discriminator(*a, **k)
        D.load_state_dict({**PD, **BD})
        made["D"] = D
        return D
    B, steps, seed0 = KC["B"], KC["steps"], KC["seed0"]
    batches = [T.kc_batch(B, seed0 + i) for i in range(steps)]
    draws, noise = [], []
    for b in batches:
        draws += [b[2], b[3].long()]
        noise += b[4]
    it, ni = iter(draws), iter(noise)
    cfg = dict(rcfg, cuda="cpu", epochs=1, batch_size=B, scaler=None, out_dir=tmp,
               generator_path=os.path.join(tmp, "g_kc.pt"))
    X = torch.cat([b[0] for b in batches]).numpy()
    y = torch.cat([b[1] for b in batches]).numpy()
    y[:4] = [0, 1, 2, 3]
    with mock.patch.object(trainer, "Discriminator", make_d), \
            mock.patch.object(trainer, "DataLoader", lambda *a, **k: _Loader([(b[0], b[1]) for b in batches])), \
            mock.patch.object(torch, "randint", lambda *a, **k: next(it).clone()), \
            mock.patch.object(torch.Tensor, "exponential_", lambda self, *a, **k: self.copy_(next(ni))):
        trainer.train_countergan(G, cfg, X, y, C.eval())
    _save("kc", [B, steps, seed0], torch.load(cfg["generator_path"]), made["D"].state_dict())


if __name__ == "__main__":
    torch.set_num_threads(4)
    with tempfile.TemporaryDirectory() as tmp:
        moons(tmp)
        kc(tmp)
