#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference,
matplotlib stubbed) on seeded synthetic inputs.  Run from the repository root in the build container:

    python tests/golden/make_golden.py

Parameters and inputs are NOT stored: they are regenerated bit-exactly from integer seeds by
``oracle.mnist_countergan.synth_params`` / ``synth_batch`` (numpy PCG64 / torch CPU generator streams),
so the fixtures stay small.  What is stored is what the reference produced:
  * forward: full ``raw`` / ``masked`` residuals, D logits, C logits, BN running stats after one
    train-mode forward, and the eval-mode residual;
  * training: after N iterations of the reference's own ``train_countergan`` (random draws injected),
    for every G / D parameter tensor and BN buffer: sum, sum of absolute values and 32 fixed samples.
"""
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mnist_countergan as O  # noqa: E402
from tests._refload import experiment  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {
    "small": dict(ch=16, nres=2, B=8, n_steps=3),
    "full": dict(ch=64, nres=6, B=8, n_steps=2),
}


def summary(t):
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, 32).long()
    return np.concatenate([[t.sum().item(), t.abs().sum().item()], t[idx].numpy()])


def build(imp, ch, nres):
    gen, dis, cls = imp("models.generator"), imp("models.discriminator"), imp("models.classifier")
    G = gen.ResidualGenerator(base_ch=ch, n_resblocks=nres)
    D = dis.Discriminator()
    C = cls.CNNClassifier().eval()
    G.load_state_dict({**O.synth_params(O.g_param_shapes(ch, nres), 1, "G"), **O.g_buffers(ch, nres)})
    D.load_state_dict(O.synth_params(O.d_param_shapes(), 2, "D"))
    C.load_state_dict(O.synth_params(O.c_param_shapes(), 3, "C"))
    for p in C.parameters():
        p.requires_grad = False
    return G, D, C


def main():
    torch.set_num_threads(8)
    with experiment("conditional_counteRGAN/mnist") as imp:
        trainer = imp("trainer")
        for name, c in CASES.items():
            ch, nres, B, n_steps = c["ch"], c["nres"], c["B"], c["n_steps"]
            out = {"meta": np.array([ch, nres, B, n_steps])}
            # ---- forward goldens
            G, D, C = build(imp, ch, nres)
            x, y, t, mask = O.synth_batch(B, 700)
            with torch.no_grad():
                raw, masked = G(x, t, mask)
                out["fwd_raw"], out["fwd_masked"] = raw.numpy(), masked.numpy()
                out["fwd_d_logits"] = D(x, y).numpy()
                out["fwd_c_logits"] = C(x).numpy()
                for k, v in G.named_buffers():
                    if "num_batches" not in k:
                        out["fwd_buf/" + k] = v.numpy().copy()
                G.eval()
                out["fwd_raw_eval"] = G(x, t, mask)[0].numpy()
            # ---- training goldens: the reference's own train_countergan, draws injected
            G, D, C = build(imp, ch, nres)
            batches = [O.synth_batch(B, 800 + i, mnist_like=(i % 2 == 1)) for i in range(n_steps)]
            it = {"i": 0}
            real_randint = torch.randint
            torch.randint = lambda *a, **k: batches[it["i"]][2].clone()

            def fake_mask(x_, ps, device, n=None):
                m = batches[it["i"]][3].clone()
                it["i"] += 1
                return m
            trainer.build_mask = fake_mask
            cfg = types.SimpleNamespace(g_lr=5e-5, d_lr=1e-5, num_epochs_gan=1, num_classes=10, patch_size=7,
                                        num_modifiable_patches=10, lambda_adv=1.0, lambda_cls=1.0, lambda_reg=2.5,
                                        lambda_mask=2.0, save_dir="/tmp", generator_path=f"/tmp/golden_{name}.pt")
            try:
                trainer.train_countergan(G, D, C, [(b[0], b[1]) for b in batches], cfg, "cpu")
            finally:
                torch.randint = real_randint
            for k, v in G.state_dict().items():
                out["train_G/" + k] = summary(v)
            for k, v in D.state_dict().items():
                out["train_D/" + k] = summary(v)
            np.savez_compressed(os.path.join(OUT, f"mnist_{name}.npz"), **out)
            print("wrote", name, sum(v.nbytes for v in out.values()), "bytes")


if __name__ == "__main__":
    main()
