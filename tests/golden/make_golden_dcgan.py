#!/usr/bin/env python
"""Generates tests/golden/dcgan.npz by running the reference script's own Generator / Discriminator classes
(dconv_gan/mnist/mnist_dcgan.py:72-116, AST-lifted from /root/reference: the script trains at import time) and its loop
body (:147-175, executed statement by statement as written, the noise draw of :156 injected) on CPU.  Run from the
repository root in the build container:

    python tests/golden/make_golden_dcgan.py

Inputs and initial parameters are NOT stored (``oracle.dcgan.synth_params`` / ``synth_batch`` regenerate them bit-exactly
from integer seeds).  Stored: errD / errG of every iteration and, for every tensor of both state_dicts after the run,
[sum, sum of absolute values, 32 fixed samples]."""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dcgan as O  # noqa: E402
from tests._refload import lift  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
B, STEPS, SEED0 = 4, 2, 11


def summary(t):
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, 32).long()
    return np.concatenate([[t.sum().item(), t.abs().sum().item()], t[idx].numpy()])


def main():
    cfg = {'batch_size': B, 'image_channel': 1, 'z_dim': 100, 'g_hidden': 64, 'd_hidden': 64, 'x_dim': 64,
           'real_label': 1., 'fake_label': 0., 'lr': 2e-4}
    ns, _ = lift("dconv_gan/mnist/mnist_dcgan.py", ("Generator", "Discriminator", "weights_init"), extra={"config": cfg})
    netG, netD = ns["Generator"](), ns["Discriminator"]()
    netG.load_state_dict({**O.synth_params(O.g_shapes(), 5), **O.buffers(O.g_shapes())})
    netD.load_state_dict({**O.synth_params(O.d_shapes(), 6), **O.buffers(O.d_shapes())})
    criterion = nn.BCELoss()
    optD = optim.Adam(netD.parameters(), lr=cfg['lr'], betas=(0.5, 0.999))
    optG = optim.Adam(netG.parameters(), lr=cfg['lr'], betas=(0.5, 0.999))
    errs = []
    for it in range(STEPS):
        real, noise = O.synth_batch(B, SEED0 + it)
        netD.zero_grad()
        label = torch.full((B,), cfg['real_label'], dtype=torch.float)
        output = netD(real)
        errD_real = criterion(output, label)
        errD_real.backward()
        fake = netG(noise)
        label.fill_(cfg['fake_label'])
        output = netD(fake.detach())
        errD_fake = criterion(output, label)
        errD_fake.backward()
        errD = errD_real + errD_fake
        optD.step()
        netG.zero_grad()
        label.fill_(cfg['real_label'])
        output = netD(fake)
        errG = criterion(output, label)
        errG.backward()
        optG.step()
        errs.append([errD.item(), errG.item()])
    out = OrderedDict(meta=np.array([B, STEPS, SEED0]), errs=np.array(errs, dtype=np.float64))
    for pre, net in (("G.", netG), ("D.", netD)):
        for k, v in net.state_dict().items():
            out[pre + k] = summary(v)
    np.savez_compressed(os.path.join(OUT, "dcgan.npz"), **out)
    print("dcgan", errs, len(out) - 2, "tensors")


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
