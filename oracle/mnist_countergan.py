"""Oracle (CPU restatement) of the MNIST conv CounteRGAN iteration.  TEST INFRASTRUCTURE.

Follows, line by line:
  * ``conditional_counteRGAN/mnist/models/generator.py:5-86``   (ResidualGenerator, _ResBlock)
  * ``conditional_counteRGAN/mnist/models/discriminator.py:5-38`` (Discriminator)
  * ``conditional_counteRGAN/mnist/models/classifier.py:4-28``   (CNNClassifier, eval mode)
  * ``conditional_counteRGAN/mnist/trainer.py:45-72``            (build_mask)
  * ``conditional_counteRGAN/mnist/trainer.py:76-132``           (one train_countergan iteration)
  * ``torch/optim/adam.py:347-547`` (single-tensor Adam, amsgrad off, wd 0)
  * ``torch/nn/functional.py`` batch_norm (train mode), binary_cross_entropy_with_logits,
    cross_entropy — published formulas.

State is a plain dict of tensors keyed by the reference's ``state_dict`` names, so a
reference module's ``state_dict()`` can be fed in directly.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

LRELU = 0.2
BN_EPS = 1e-5
BN_MOM = 0.1


# --------------------------------------------------------------------------- init
def g_param_shapes(base_ch=64, n_resblocks=6, num_classes=10, img_shape=(1, 28, 28)):
    """Parameter names/shapes in ``ResidualGenerator.parameters()`` order (generator.py:32-56)."""
    C, H, W = img_shape
    s = OrderedDict()
    s["embed.weight"] = (num_classes, H * W)
    s["conv_in.weight"] = (base_ch, C + 2, 3, 3)
    s["conv_in.bias"] = (base_ch,)
    for i in range(n_resblocks):
        p = f"resblocks.{i}."
        s[p + "conv1.weight"] = (base_ch, base_ch, 3, 3)
        s[p + "conv1.bias"] = (base_ch,)
        s[p + "bn1.weight"] = (base_ch,)
        s[p + "bn1.bias"] = (base_ch,)
        s[p + "conv2.weight"] = (base_ch, base_ch, 3, 3)
        s[p + "conv2.bias"] = (base_ch,)
        s[p + "bn2.weight"] = (base_ch,)
        s[p + "bn2.bias"] = (base_ch,)
    s["conv_mid.weight"] = (base_ch, base_ch, 3, 3)
    s["conv_mid.bias"] = (base_ch,)
    s["conv_out.weight"] = (1, base_ch, 3, 3)
    s["conv_out.bias"] = (1,)
    return s


def d_param_shapes(num_classes=10, img_shape=(1, 28, 28), d_hidden=64):
    """``Discriminator.parameters()`` order (discriminator.py:9-31)."""
    C, H, W = img_shape
    s = OrderedDict()
    s["cond_embed.weight"] = (num_classes, H * W)
    s["main.0.weight"] = (d_hidden, 2, 3, 3)
    s["main.2.weight"] = (d_hidden * 2, d_hidden, 3, 3)
    s["main.4.weight"] = (d_hidden * 4, d_hidden * 2, 3, 3)
    s["main.6.weight"] = (d_hidden * 4, d_hidden * 4, 3, 3)
    s["adv_head.weight"] = (1, d_hidden * 4)
    s["adv_head.bias"] = (1,)
    return s


def c_param_shapes(num_classes=10):
    """``CNNClassifier.parameters()`` order (classifier.py:7-22)."""
    s = OrderedDict()
    s["conv.0.weight"] = (32, 1, 3, 3)
    s["conv.0.bias"] = (32,)
    s["conv.2.weight"] = (64, 32, 3, 3)
    s["conv.2.bias"] = (64,)
    s["conv.4.weight"] = (128, 64, 3, 3)
    s["conv.4.bias"] = (128,)
    s["fc.1.weight"] = (256, 128 * 7 * 7)
    s["fc.1.bias"] = (256,)
    s["fc.4.weight"] = (num_classes, 256)
    s["fc.4.bias"] = (num_classes,)
    return s


def synth_params(shapes, seed, kind):
    """Torch-independent synthetic parameters (numpy PCG64, stable across versions).

    Scales mimic the reference initialisers (generator.py:58-69 Kaiming a=0.2 fan_in,
    embedding N(0, .01); torch defaults elsewhere) closely enough to be well conditioned;
    the exact distribution is irrelevant for parity because the same tensors feed both sides.
    BN affine parameters are perturbed away from (1, 0) so their gradients are exercised.
    """
    import numpy as np

    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shp in shapes.items():
        n = int(np.prod(shp))
        if "embed" in name:
            a = rng.standard_normal(n) * (0.01 if kind == "G" else 1.0)
        elif ".bn" in name and name.endswith("weight"):
            a = 1.0 + 0.1 * rng.standard_normal(n)
        elif ".bn" in name and name.endswith("bias"):
            a = 0.1 * rng.standard_normal(n)
        elif name.endswith("bias"):
            a = 0.05 * rng.standard_normal(n)
        else:
            fan_in = int(np.prod(shp[1:]))
            a = rng.standard_normal(n) * math.sqrt(2.0 / (1 + LRELU ** 2) / fan_in)
        out[name] = torch.from_numpy(a.astype("float32").reshape(shp))
    return out


def g_buffers(base_ch=64, n_resblocks=6):
    b = OrderedDict()
    for i in range(n_resblocks):
        for j in (1, 2):
            p = f"resblocks.{i}.bn{j}."
            b[p + "running_mean"] = torch.zeros(base_ch)
            b[p + "running_var"] = torch.ones(base_ch)
            b[p + "num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return b


def is_bn_shadowed_bias(name):
    """Conv biases that feed straight into a train-mode BatchNorm (generator.py:11-15): their true
    gradient is exactly zero (BN subtracts the batch mean), so what the reference's Adam sees is
    rounding noise of magnitude ~1e-9, which Adam (eps 1e-8) turns into a random +-lr walk.
    Parity on these tensors is therefore bounded by lr*steps, not by a relative tolerance."""
    return name.startswith("resblocks.") and (name.endswith("conv1.bias") or name.endswith("conv2.bias"))


# --------------------------------------------------------------------------- ops
def batch_norm_train(x, gamma, beta, running_mean, running_var, nbt, update=True):
    """Train-mode BatchNorm2d: normalise with the *biased* batch variance, update the
    running stats with the *unbiased* one, momentum 0.1, eps 1e-5 (generator.py:12,15)."""
    n = x.numel() // x.shape[1]
    mean = x.mean(dim=(0, 2, 3))
    var = ((x - mean[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
    if update:
        with torch.no_grad():
            running_mean.mul_(1 - BN_MOM).add_(BN_MOM * mean.detach())
            running_var.mul_(1 - BN_MOM).add_(BN_MOM * var.detach() * (n / max(n - 1, 1)))
            nbt.add_(1)
    xhat = (x - mean[None, :, None, None]) * torch.rsqrt(var + BN_EPS)[None, :, None, None]
    return xhat * gamma[None, :, None, None] + beta[None, :, None, None]


def batch_norm_eval(x, gamma, beta, running_mean, running_var):
    s = gamma * torch.rsqrt(running_var + BN_EPS)
    return x * s[None, :, None, None] + (beta - running_mean * s)[None, :, None, None]


def bce_with_logits(z, t):
    """mean( max(z,0) - z*t + log1p(exp(-|z|)) ) — nn.BCEWithLogitsLoss (trainer.py:79)."""
    return (z.clamp_min(0) - z * t + torch.log1p(torch.exp(-z.abs()))).mean()


def cross_entropy(logits, target):
    """mean( logsumexp(l) - l[target] ) — nn.CrossEntropyLoss (trainer.py:80)."""
    lse = torch.logsumexp(logits, dim=1)
    return (lse - logits.gather(1, target[:, None]).squeeze(1)).mean()


def build_mask_from_patches(patch_idx, bs, h=28, w=28, patch_size=7):
    """Deterministic half of build_mask (trainer.py:51-72): ``patch_idx[b]`` lists the
    modifiable patches of sample b (row-major over the 4x4 patch grid); nearest upsample."""
    nh, nw = h // patch_size, w // patch_size
    pm = torch.zeros(bs, 1, nh, nw)
    for b in range(bs):
        pm.view(bs, -1)[b, patch_idx[b]] = 1.0
    return F.interpolate(pm, size=(h, w), mode="nearest")


def random_patch_idx(bs, gen, total=16, k=10):
    return [torch.randperm(total, generator=gen)[:k] for _ in range(bs)]


# --------------------------------------------------------------------------- bf16 noise floor
class _RoundBoth(torch.autograd.Function):
    """Rounds a tensor to bf16 in the forward pass and its gradient to bf16 in the backward pass."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().to(g.dtype)


import contextlib


@contextlib.contextmanager
def bf16_storage_emulation():
    """Within this context the oracle rounds every conv input/output and activation output (and the
    matching gradients) to bf16, and the weights of the wide (>=32x32 channel) convolutions, i.e. it
    emulates a bf16-storage / fp32-accumulate execution of the SAME algorithm.  The distance between
    this run and the plain fp32 run is the *noise floor* any bf16 implementation of the step has; the
    GPU tests bound the native bf16 path by a small multiple of it instead of by a hand-picked number."""
    rb = _RoundBoth.apply
    oc, ol, orl = F.conv2d, F.leaky_relu, F.relu

    def conv2d(inp, w, b=None, **kw):
        wide = w.shape[0] >= 32 and w.shape[1] >= 32
        return rb(oc(rb(inp), rb(w) if wide else w, b, **kw))

    F.conv2d = conv2d
    F.leaky_relu = lambda v, s=0.01: rb(ol(v, s))
    F.relu = lambda v: rb(orl(v))
    try:
        yield
    finally:
        F.conv2d, F.leaky_relu, F.relu = oc, ol, orl


# --------------------------------------------------------------------------- nets
def g_forward(P, Bf, x, target, mask, n_resblocks=6, residual_scaling=0.1, training=True,
              taps=None):
    """ResidualGenerator.forward (generator.py:71-86)."""
    B, C, H, W = x.shape
    y_map = P["embed.weight"][target].view(B, 1, H, W)
    inp = torch.cat([x, y_map, mask], dim=1)
    h = F.leaky_relu(F.conv2d(inp, P["conv_in.weight"], P["conv_in.bias"], padding=1), LRELU)
    if taps is not None:
        taps["h0"] = h
    for i in range(n_resblocks):
        p = f"resblocks.{i}."

        def bn(t, j):
            q = p + f"bn{j}."
            if training:
                return batch_norm_train(t, P[q + "weight"], P[q + "bias"], Bf[q + "running_mean"],
                                        Bf[q + "running_var"], Bf[q + "num_batches_tracked"])
            return batch_norm_eval(t, P[q + "weight"], P[q + "bias"], Bf[q + "running_mean"],
                                   Bf[q + "running_var"])

        y1 = F.conv2d(h, P[p + "conv1.weight"], P[p + "conv1.bias"], padding=1)
        z1 = F.leaky_relu(bn(y1, 1), LRELU)
        y2 = F.conv2d(z1, P[p + "conv2.weight"], P[p + "conv2.bias"], padding=1)
        h = h + 0.1 * bn(y2, 2)          # generator.py:22 (the 0.1 is hard-coded there)
        if taps is not None:
            taps[f"y1.{i}"], taps[f"z1.{i}"], taps[f"y2.{i}"], taps[f"h.{i + 1}"] = y1, z1, y2, h
    h = F.leaky_relu(F.conv2d(h, P["conv_mid.weight"], P["conv_mid.bias"], padding=1), LRELU)
    if taps is not None:
        taps["hm"] = h
    raw = F.conv2d(h, P["conv_out.weight"], P["conv_out.bias"], padding=1) * residual_scaling
    masked = raw * mask
    return raw, masked


def d_forward(P, x, cond_idx, taps=None):
    """Discriminator.forward (discriminator.py:33-38)."""
    B, C, H, W = x.shape
    cond_map = P["cond_embed.weight"][cond_idx].view(B, 1, H, W)
    z = torch.cat([x, cond_map], dim=1)
    for k in ("main.0", "main.2", "main.4", "main.6"):
        z = F.leaky_relu(F.conv2d(z, P[k + ".weight"], None, stride=2, padding=1), LRELU)
        if taps is not None:
            taps[k] = z
    f = z.mean(dim=(2, 3))
    return f @ P["adv_head.weight"].t() + P["adv_head.bias"]


def c_forward(P, x, taps=None):
    """CNNClassifier.forward in eval mode (classifier.py:25-28; dropout = identity, main.py:31)."""
    z = F.relu(F.conv2d(x, P["conv.0.weight"], P["conv.0.bias"], stride=1, padding=1))
    z = F.relu(F.conv2d(z, P["conv.2.weight"], P["conv.2.bias"], stride=2, padding=1))
    z = F.relu(F.conv2d(z, P["conv.4.weight"], P["conv.4.bias"], stride=2, padding=1))
    z = z.flatten(1)
    z = F.relu(z @ P["fc.1.weight"].t() + P["fc.1.bias"])
    return z @ P["fc.4.weight"].t() + P["fc.4.bias"]


# --------------------------------------------------------------------------- Adam
def adam_init(P):
    return {"step": 0,
            "exp_avg": OrderedDict((k, torch.zeros_like(v)) for k, v in P.items()),
            "exp_avg_sq": OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())}


def adam_step(P, G, A, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor path (torch/optim/adam.py:347-547), defaults of
    trainer.py:77-78.  Parameters whose grad is None are skipped, as torch does."""
    A["step"] += 1
    t = A["step"]
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    with torch.no_grad():
        for k, p in P.items():
            g = G.get(k)
            if g is None:
                continue
            m, v = A["exp_avg"][k], A["exp_avg_sq"][k]
            m.lerp_(g, 1 - beta1)
            v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
            denom = (v.sqrt() / bc2_sqrt).add_(eps)
            p.addcdiv_(m, denom, value=-step_size)


# --------------------------------------------------------------------------- the step
class Hyper:
    """Defaults of conditional_counteRGAN/mnist/config.py:10-15."""
    g_lr = 5e-5
    d_lr = 1e-5
    lambda_adv = 1.0
    lambda_cls = 1.0
    lambda_reg = 2.5
    lambda_mask = 2.0

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


def make_state(PG, BG, PD, PC, dtype=torch.float32):
    """Deep-copies parameter dicts into an oracle state (leaf tensors, Adam zeros)."""
    def cp(d, grad):
        o = OrderedDict()
        for k, v in d.items():
            t = v.detach().clone()
            if t.is_floating_point():
                t = t.to(dtype)
                t.requires_grad_(grad)
            o[k] = t
        return o
    S = {"G": cp(PG, True), "GB": cp(BG, False), "D": cp(PD, True), "C": cp(PC, False)}
    S["adam_g"] = adam_init(S["G"])
    S["adam_d"] = adam_init(S["D"])
    return S


def countergan_step(S, x, y, target, mask, hp=None, n_resblocks=6, pollute_d=False, taps=None):
    """One iteration of train_countergan (trainer.py:89-132) with the random draws
    (target_y :94, mask :95) injected.  Mutates ``S`` in place.  Returns (scalars, grads)."""
    hp = hp or Hyper()
    G, GB, D, C = S["G"], S["GB"], S["D"], S["C"]
    dt = next(iter(G.values())).dtype
    x = x.to(dt)
    mask = mask.to(dt)

    raw, masked = g_forward(G, GB, x, target, mask, n_resblocks, taps=taps)      # :96
    x_cf = torch.clamp(x + masked, -1.0, 1.0)                                     # :97
    mask_pen = torch.mean(torch.abs(raw * (1.0 - mask)))                          # :99

    # --- D update (:102-112)
    d_real = d_forward(D, x, y)
    d_fake = d_forward(D, x_cf.detach(), target)
    d_loss = bce_with_logits(d_real, torch.ones_like(d_real)) + \
        bce_with_logits(d_fake, torch.zeros_like(d_fake))
    gD = dict(zip(D.keys(), torch.autograd.grad(d_loss, list(D.values()))))
    adam_step(D, gD, S["adam_d"], hp.d_lr)

    # --- G update (:115-123): D here is the *updated* D
    g_fake = d_forward(D, x_cf, target)
    g_adv = bce_with_logits(g_fake, torch.ones_like(g_fake))
    g_cls = cross_entropy(c_forward(C, x_cf), target)
    reg_l1 = masked.abs().mean()
    g_loss = hp.lambda_adv * g_adv + hp.lambda_cls * g_cls + hp.lambda_reg * reg_l1 + \
        hp.lambda_mask * mask_pen
    wrt = list(G.values()) + (list(D.values()) if pollute_d else [])
    gr = torch.autograd.grad(g_loss, wrt, allow_unused=True)
    gG = dict(zip(G.keys(), gr[:len(G)]))
    gD_g = dict(zip(D.keys(), gr[len(G):])) if pollute_d else None
    adam_step(G, gG, S["adam_g"], hp.g_lr)

    with torch.no_grad():
        sc = {
            "d_loss": d_loss.item(), "g_loss": g_loss.item(), "g_adv": g_adv.item(),
            "g_cls": g_cls.item(), "reg_l1": reg_l1.item(), "mask_pen": mask_pen.item(),
            "d_real_p": torch.sigmoid(d_real).mean().item(),
            "d_fake_p": torch.sigmoid(d_fake).mean().item(),
        }
    grads = {"D": gD, "G": gG, "D_from_g": gD_g,
             "x_cf": x_cf.detach(), "raw": raw.detach(), "masked": masked.detach(),
             "d_real": d_real.detach(), "d_fake": d_fake.detach(), "g_fake": g_fake.detach()}
    return sc, grads


def synth_batch(B, seed, mnist_like=False):
    """Seeded synthetic batch of SURVEY.md §8d: x~U(-1,1) (optionally 80 % of pixels
    exactly -1 to exercise the clamp edge), labels/targets uniform, 10-of-16 patch masks."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 1, 28, 28, generator=g) * 2 - 1
    if mnist_like:
        x = torch.where(torch.rand(B, 1, 28, 28, generator=g) < 0.8, torch.full_like(x, -1.0), x)
    y = torch.randint(0, 10, (B,), generator=g)
    t = torch.randint(0, 10, (B,), generator=g)
    mask = build_mask_from_patches(random_patch_idx(B, g), B)
    return x, y, t, mask
