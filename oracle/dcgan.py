"""Oracle (CPU restatement) of the DCGAN iteration.  TEST INFRASTRUCTURE.

Follows ``dconv_gan/mnist/mnist_dcgan.py``: Generator :72-93, Discriminator :96-116, weights_init :63-69,
loop body :147-175 (noise at :156 injected).  Parameters / buffers are dicts keyed by the reference's
``state_dict`` names (``main.<i>.weight`` ...).  The network works on 1x64x64 images (the 28x28 source is
resized at :43); z is [B,100,1,1].
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

from .mnist_countergan import adam_init, adam_step, batch_norm_train

G_CH = [100, 512, 256, 128, 64, 1]
D_CH = [1, 64, 128, 256, 512, 1]


def g_shapes():
    s = OrderedDict()
    for i in range(5):
        s[f"main.{3 * i}.weight"] = (G_CH[i], G_CH[i + 1], 4, 4)        # ConvTranspose2d: [in, out, k, k]
        if i < 4:
            s[f"main.{3 * i + 1}.weight"] = (G_CH[i + 1],)
            s[f"main.{3 * i + 1}.bias"] = (G_CH[i + 1],)
    return s


def d_shapes():
    s = OrderedDict()
    s["main.0.weight"] = (64, 1, 4, 4)
    idx = 2
    for i in range(1, 4):
        s[f"main.{idx}.weight"] = (D_CH[i + 1], D_CH[i], 4, 4)
        s[f"main.{idx + 1}.weight"] = (D_CH[i + 1],)
        s[f"main.{idx + 1}.bias"] = (D_CH[i + 1],)
        idx += 3
    s["main.11.weight"] = (1, 512, 4, 4)
    return s


def bn_names(shapes):
    return [k[:-7] for k, v in shapes.items() if k.endswith(".weight") and len(v) == 1]


def buffers(shapes):
    b = OrderedDict()
    for n in bn_names(shapes):
        C = shapes[n + ".weight"][0]
        b[n + ".running_mean"] = torch.zeros(C)
        b[n + ".running_var"] = torch.ones(C)
        b[n + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return b


def synth_params(shapes, seed):
    """weights_init (:63-69): conv ~ N(0, .02), BN weight ~ N(1, .02), BN bias 0 — from numpy PCG64."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for k, s in shapes.items():
        if len(s) == 4:
            a = rng.standard_normal(s) * 0.02
        elif k.endswith("weight"):
            a = 1.0 + 0.02 * rng.standard_normal(s)
        else:
            a = 0.05 * rng.standard_normal(s)       # perturbed away from 0 so the path is exercised
        out[k] = torch.from_numpy(a.astype("float32"))
    return out


def _bn(P, Bf, name, x, training):
    if training:
        return batch_norm_train(x, P[name + ".weight"], P[name + ".bias"], Bf[name + ".running_mean"],
                                Bf[name + ".running_var"], Bf[name + ".num_batches_tracked"])
    s = P[name + ".weight"] * torch.rsqrt(Bf[name + ".running_var"] + 1e-5)
    return x * s[None, :, None, None] + (P[name + ".bias"] - Bf[name + ".running_mean"] * s)[None, :, None, None]


def g_forward(P, Bf, z, training=True, taps=None):
    x = z
    for i in range(5):
        stride, pad = (1, 0) if i == 0 else (2, 1)
        x = F.conv_transpose2d(x, P[f"main.{3 * i}.weight"], None, stride, pad)
        if i < 4:
            x = torch.relu(_bn(P, Bf, f"main.{3 * i + 1}", x, training))
        if taps is not None:
            taps[f"g{i}"] = x
    return torch.tanh(x)


def d_forward(P, Bf, x, training=True, taps=None):
    x = F.leaky_relu(F.conv2d(x, P["main.0.weight"], None, 2, 1), 0.2)
    idx = 2
    for i in range(1, 4):
        x = F.conv2d(x, P[f"main.{idx}.weight"], None, 2, 1)
        x = F.leaky_relu(_bn(P, Bf, f"main.{idx + 1}", x, training), 0.2)
        if taps is not None:
            taps[f"d{i}"] = x
        idx += 3
    x = F.conv2d(x, P["main.11.weight"], None, 1, 0)
    return torch.sigmoid(x).view(-1, 1).squeeze(1)


def bce(p, t):
    """nn.BCELoss: -mean(t log p + (1-t) log(1-p)), logs clamped at -100."""
    return -(t * torch.log(p).clamp_min(-100) + (1 - t) * torch.log(1 - p).clamp_min(-100)).mean()


def make_state(PG, BG, PD, BD):
    def cp(d, grad):
        o = OrderedDict()
        for k, v in d.items():
            t = v.detach().clone()
            if t.is_floating_point():
                t.requires_grad_(grad)
            o[k] = t
        return o
    S = {"G": cp(PG, True), "GB": cp(BG, False), "D": cp(PD, True), "DB": cp(BD, False)}
    S["adam_g"], S["adam_d"] = adam_init(S["G"]), adam_init(S["D"])
    return S


def dcgan_step(S, real, noise, lr=2e-4, betas=(0.5, 0.999)):
    """mnist_dcgan.py:147-175."""
    G, GB, D, DB = S["G"], S["GB"], S["D"], S["DB"]
    ones, zeros = torch.ones(real.shape[0]), torch.zeros(real.shape[0])
    out_real = d_forward(D, DB, real)
    errD_real = bce(out_real, ones)
    g1 = torch.autograd.grad(errD_real, list(D.values()))
    fake = g_forward(G, GB, noise)
    out_fake = d_forward(D, DB, fake.detach())
    errD_fake = bce(out_fake, zeros)
    g2 = torch.autograd.grad(errD_fake, list(D.values()))
    gD = {k: a + b for k, a, b in zip(D.keys(), g1, g2)}
    adam_step(D, gD, S["adam_d"], lr, betas[0], betas[1])
    out2 = d_forward(D, DB, fake)
    errG = bce(out2, ones)
    gG = dict(zip(G.keys(), torch.autograd.grad(errG, list(G.values()))))
    adam_step(G, gG, S["adam_g"], lr, betas[0], betas[1])
    sc = {"errD": (errD_real + errD_fake).item(), "errG": errG.item(), "D_x": out_real.mean().item(),
          "D_G_z1": out_fake.mean().item(), "D_G_z2": out2.mean().item()}
    return sc, {"D": gD, "G": gG, "fake": fake.detach()}


def synth_batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    real28 = torch.rand(B, 1, 28, 28, generator=g) * 2 - 1
    real = F.interpolate(real28, size=(64, 64), mode="bilinear", align_corners=False)   # transforms.Resize(64), :43
    noise = torch.randn(B, 100, 1, 1, generator=g)
    return real, noise
