"""Oracle (CPU restatement) of the two MLP GAN iterations on 2-D points.  TEST INFRASTRUCTURE.

Follows
  * ``conditional_gan/moons/make_moons_cgan.py:35-60`` (Generator / Discriminator) and ``:90-129`` (loop body)
  * ``simple_gan/moons/make_moons_gan.py:33-46`` (build_generator / build_discriminator) and ``:61-88`` (loop body)
with the random draws (z at :63/:75 resp. :97/:116, fake labels at :98/:117) injected.

Parameters are dicts keyed like the reference's ``state_dict`` with the ``net.`` prefix dropped:
``0.weight [H, in]``, ``0.bias``, ``2.weight [out, H]``, ``2.bias``.
"""
import math
from collections import OrderedDict

import torch

from .mnist_countergan import adam_init, adam_step


def shapes(z_dim=32, label_dim=2, hidden=128):
    g = OrderedDict([("0.weight", (hidden, z_dim + label_dim)), ("0.bias", (hidden,)), ("2.weight", (2, hidden)),
                     ("2.bias", (2,))])
    d = OrderedDict([("0.weight", (hidden, 2 + label_dim)), ("0.bias", (hidden,)), ("2.weight", (1, hidden)),
                     ("2.bias", (1,))])
    return g, d


def synth_params(shp, seed):
    """nn.Linear-like U(-1/sqrt(in), 1/sqrt(in)) from numpy PCG64 (torch independent)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    fan = None
    for k, s in shp.items():
        if k.endswith("weight"):
            fan = s[1]
        b = 1.0 / math.sqrt(fan)
        out[k] = torch.from_numpy(rng.uniform(-b, b, size=s).astype("float32"))
    return out


def mlp(P, x):
    h = torch.relu(x @ P["0.weight"].t() + P["0.bias"])
    return h @ P["2.weight"].t() + P["2.bias"]


def g_forward(P, z, onehot=None):
    """Generator.forward (make_moons_cgan.py:44-46) / the Sequential of make_moons_gan.py:33-38."""
    return mlp(P, z if onehot is None else torch.cat([z, onehot], 1))


def d_forward(P, x, onehot=None):
    """Discriminator.forward incl. the final Sigmoid (make_moons_cgan.py:58-60, make_moons_gan.py:40-46)."""
    return torch.sigmoid(mlp(P, x if onehot is None else torch.cat([x, onehot], 1)))


def make_state(PG, PD):
    cp = lambda d: OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in d.items())  # noqa: E731
    S = {"G": cp(PG), "D": cp(PD)}
    S["adam_g"], S["adam_d"] = adam_init(S["G"]), adam_init(S["D"])
    return S


def gan_step(S, real, real_oh, z1, oh1, z2, oh2, lr=1e-3):
    """One iteration.  ``*_oh`` are None for the unconditional GAN.  Returns (loss_D, loss_G, grads)."""
    G, D = S["G"], S["D"]
    fake = g_forward(G, z1, oh1)
    # the conditional script detaches (cgan :104); the simple one does not (gan :64-72) but only D is stepped and
    # G's gradients are zeroed before its own step, so detaching is equivalent for every quantity that survives
    D_real = d_forward(D, real, real_oh)
    D_fake = d_forward(D, fake.detach(), oh1)
    loss_D = -torch.mean(torch.log(D_real) + torch.log(1 - D_fake))
    gD = dict(zip(D.keys(), torch.autograd.grad(loss_D, list(D.values()))))
    adam_step(D, gD, S["adam_d"], lr)
    fake2 = g_forward(G, z2, oh2)
    D_fake2 = d_forward(D, fake2, oh2)
    loss_G = -torch.mean(torch.log(D_fake2))
    gG = dict(zip(G.keys(), torch.autograd.grad(loss_G, list(G.values()))))
    adam_step(G, gG, S["adam_g"], lr)
    return loss_D.item(), loss_G.item(), {"D": gD, "G": gG, "fake": fake.detach(), "fake2": fake2.detach()}


def synth_batch(B, seed, z_dim=32, label_dim=2):
    g = torch.Generator().manual_seed(seed)
    t = torch.rand(B, generator=g) * math.pi
    lab = torch.randint(0, 2, (B,), generator=g)
    real = torch.stack([torch.cos(t) * (1 - 2 * lab) + lab, torch.sin(t) * (1 - 2 * lab) + 0.5 * lab], 1)
    real = real + 0.05 * torch.randn(B, 2, generator=g)
    z1, z2 = torch.randn(B, z_dim, generator=g), torch.randn(B, z_dim, generator=g)
    if label_dim == 0:
        return real, None, z1, None, z2, None
    oh = lambda l: torch.nn.functional.one_hot(l, label_dim).float()  # noqa: E731
    lab1 = torch.zeros(B, dtype=torch.long)                    # randint(0, 1) of cgan :98 -> always class 0
    lab2 = torch.randint(0, label_dim, (B,), generator=g)
    return real, oh(lab), z1, oh(lab1), z2, oh(lab2)
