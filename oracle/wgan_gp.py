"""Oracle (CPU restatement) of the conditional WGAN-GP iteration.  TEST INFRASTRUCTURE.

Follows ``conditional_gan/mnist/mnist_wgan_conditional.py``: Hyperparameter :20-31, Generator :51-78, Critic :80-108,
optimisers :116-117 (AdamW, lr 1e-4, betas (0, 0.9), torch's default weight decay 1e-2), critic update :132-154 (gradient
penalty :146-150 through ``autograd.grad(create_graph=True)``), generator update every ``n_critic`` batches :156-168.
The draws of the loop (noise :139/:161, alpha :144, the generator's labels :160) are explicit inputs.  Parameters /
buffers are dicts keyed by the reference's ``state_dict`` names.  Pinned against the AST-lifted reference classes and
loop body by ``tests/test_wgan_oracle_vs_reference.py`` and the committed ``tests/golden/wgan_gp.npz``.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

from .mnist_countergan import adam_init, adam_step, batch_norm_train


class Hyper:
    """mnist_wgan_conditional.py:20-31."""

    def __init__(self, num_classes=10, batchsize=128, latent_size=32, n_critic=5, critic_size=1024, generator_size=1024,
                 critic_hidden_size=1024, gp_lambda=10.0, lr=1e-4, betas=(0.0, 0.9), weight_decay=1e-2):
        self.num_classes, self.batchsize, self.latent_size, self.n_critic = num_classes, batchsize, latent_size, n_critic
        self.critic_size, self.generator_size, self.critic_hidden_size = critic_size, generator_size, critic_hidden_size
        self.gp_lambda, self.lr, self.betas, self.weight_decay = gp_lambda, lr, betas, weight_decay


def g_shapes(hp):
    g = hp.generator_size
    s = OrderedDict()
    s["latent_embedding.0.weight"], s["latent_embedding.0.bias"] = (g // 2, hp.latent_size), (g // 2,)
    s["condition_embedding.0.weight"], s["condition_embedding.0.bias"] = (g // 2, hp.num_classes), (g // 2,)
    chans = [(g, g, 4), (g, g // 2, 3), (g // 2, g // 4, 4), (g // 4, 1, 4)]          # ConvTranspose2d: [in, out, k, k]
    for i, (ci, co, k) in enumerate(chans):
        s[f"tcnn.{3 * i}.weight"], s[f"tcnn.{3 * i}.bias"] = (ci, co, k, k), (co,)
        if i < 3:
            s[f"tcnn.{3 * i + 1}.weight"], s[f"tcnn.{3 * i + 1}.bias"] = (co,), (co,)
    return s


def c_shapes(hp):
    c = hp.critic_size
    s = OrderedDict()
    s["condition_embedding.0.weight"], s["condition_embedding.0.bias"] = (c * 4, hp.num_classes), (c * 4,)
    for i, (ci, co) in enumerate([(1, c // 4), (c // 4, c // 2), (c // 2, c)]):
        s[f"cnn_net.{3 * i}.weight"], s[f"cnn_net.{3 * i}.bias"] = (co, ci, 3, 3), (co,)
        s[f"cnn_net.{3 * i + 1}.weight"], s[f"cnn_net.{3 * i + 1}.bias"] = (co,), (co,)     # InstanceNorm2d(affine=True)
    s["Critic_net.0.weight"], s["Critic_net.0.bias"] = (hp.critic_hidden_size, c * 8), (hp.critic_hidden_size,)
    s["Critic_net.2.weight"], s["Critic_net.2.bias"] = (1, hp.critic_hidden_size), (1,)
    return s


# conv biases directly in front of a BatchNorm / InstanceNorm: no effect on any loss, analytically zero gradient
SHADOWED = ("tcnn.0.bias", "tcnn.3.bias", "tcnn.6.bias", "cnn_net.0.bias", "cnn_net.3.bias", "cnn_net.6.bias")


def g_bn_names():
    return ["tcnn.1", "tcnn.4", "tcnn.7"]


def g_buffers(hp):
    b = OrderedDict()
    sh = g_shapes(hp)
    for n in g_bn_names():
        C = sh[n + ".weight"][0]
        b[n + ".running_mean"], b[n + ".running_var"] = torch.zeros(C), torch.ones(C)
        b[n + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return b


def synth_params(shapes, seed):
    """Fan-in scaled normal weights, norm weights around 1, small non-zero biases (numpy PCG64): torch's default
    initialisers are not restated, parity is on the update rule given the same parameters."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for k, s in shapes.items():
        if len(s) >= 2:
            fan = int(np.prod(s[1:])) if not k.startswith("tcnn") else s[0] * (1 if k == "tcnn.0.weight" else 4)
            a = rng.standard_normal(s) / math.sqrt(fan)
        elif k.endswith("weight"):
            a = 1.0 + 0.1 * rng.standard_normal(s)
        else:
            a = 0.05 * rng.standard_normal(s)
        out[k] = torch.from_numpy(a.astype("float32"))
    return out


def g_forward(P, Bf, latent, cond, training=True, taps=None):
    """Generator.forward :74-78."""
    v = torch.cat([F.linear(latent, P["latent_embedding.0.weight"], P["latent_embedding.0.bias"]),
                   F.linear(cond, P["condition_embedding.0.weight"], P["condition_embedding.0.bias"])], dim=1)
    x = v.reshape(v.shape[0], -1, 1, 1)
    geo = [(1, 0), (2, 1), (2, 1), (2, 1)]
    for i, (stride, pad) in enumerate(geo):
        x = F.conv_transpose2d(x, P[f"tcnn.{3 * i}.weight"], P[f"tcnn.{3 * i}.bias"], stride, pad)
        if i < 3:
            n = f"tcnn.{3 * i + 1}"
            if training:
                x = batch_norm_train(x, P[n + ".weight"], P[n + ".bias"], Bf[n + ".running_mean"], Bf[n + ".running_var"],
                                     Bf[n + ".num_batches_tracked"])
            else:
                s = P[n + ".weight"] * torch.rsqrt(Bf[n + ".running_var"] + 1e-5)
                x = x * s[None, :, None, None] + (P[n + ".bias"] - Bf[n + ".running_mean"] * s)[None, :, None, None]
            x = torch.relu(x)
        if taps is not None:
            taps[f"g{i}"] = x
    return torch.tanh(x)


def instance_norm(x, w, b, eps=1e-5):
    """nn.InstanceNorm2d(affine=True, track_running_stats=False): per (sample, channel) statistics, biased variance."""
    mu = x.mean((2, 3), keepdim=True)
    var = ((x - mu) ** 2).mean((2, 3), keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w[None, :, None, None] + b[None, :, None, None]


def c_forward(P, image, cond, taps=None):
    """Critic.forward :102-106."""
    vc = F.linear(cond, P["condition_embedding.0.weight"], P["condition_embedding.0.bias"])
    x = image
    for i in range(3):
        x = F.conv2d(x, P[f"cnn_net.{3 * i}.weight"], P[f"cnn_net.{3 * i}.bias"], 2)
        x = F.leaky_relu(instance_norm(x, P[f"cnn_net.{3 * i + 1}.weight"], P[f"cnn_net.{3 * i + 1}.bias"]), 0.2)
        if taps is not None:
            taps[f"c{i}"] = x
    f = torch.cat([x.flatten(1), vc], dim=1)
    h = F.leaky_relu(F.linear(f, P["Critic_net.0.weight"], P["Critic_net.0.bias"]), 0.2)
    return F.linear(h, P["Critic_net.2.weight"], P["Critic_net.2.bias"])


def adamw_step(P, G, A, lr, beta1, beta2, weight_decay, eps=1e-8):
    """torch/optim/adamw.py: p *= 1 - lr * wd, then the Adam update."""
    with torch.no_grad():
        for k, p in P.items():
            if G.get(k) is not None:
                p.mul_(1 - lr * weight_decay)
    adam_step(P, G, A, lr, beta1, beta2, eps)


def make_state(PG, BG, PC):
    def cp(d, grad):
        o = OrderedDict()
        for k, v in d.items():
            t = v.detach().clone()
            if t.is_floating_point():
                t.requires_grad_(grad)
            o[k] = t
        return o
    S = {"G": cp(PG, True), "GB": cp(BG, False), "C": cp(PC, True)}
    S["adam_g"], S["adam_c"] = adam_init(S["G"]), adam_init(S["C"])
    return S


def critic_step(S, hp, real, cond, noise, alpha, update=True):
    """:132-154.  ``alpha`` is [B,1].  Returns scalars and the critic's gradients (before the update)."""
    G, GB, C = S["G"], S["GB"], S["C"]
    out_real = c_forward(C, real, cond)
    loss_real = out_real.mean()
    with torch.no_grad():
        fake = g_forward(G, GB, noise, cond)                      # train-mode BatchNorm: the running stats move (:140)
    out_fake = c_forward(C, fake, cond)
    loss_fake = out_fake.mean()
    a4 = alpha.view(-1, 1, 1, 1)
    inter = (a4 * real + (1.0 - a4) * fake).requires_grad_(True)
    d_inter = c_forward(C, inter, cond)
    grads = torch.autograd.grad(d_inter, inter, torch.ones_like(d_inter), create_graph=True)[0]
    norms = grads.reshape(grads.shape[0], -1).norm(dim=1)
    gp = hp.gp_lambda * ((norms - 1.0) ** 2).mean()
    loss = -loss_real + loss_fake + gp
    gC = dict(zip(C.keys(), torch.autograd.grad(loss, list(C.values()), allow_unused=True)))
    gC = {k: (torch.zeros_like(C[k]) if v is None else v) for k, v in gC.items()}
    if update:
        adamw_step(C, gC, S["adam_c"], hp.lr, hp.betas[0], hp.betas[1], hp.weight_decay)
    sc = {"critic_loss": loss.item(), "loss_real": loss_real.item(), "loss_fake": loss_fake.item(), "gp": gp.item()}
    return sc, {"C": gC, "fake": fake.detach(), "grad_norms": norms.detach(), "inter_grad": grads.detach()}


def generator_step(S, hp, noise, cond, update=True):
    """:156-168."""
    G, GB, C = S["G"], S["GB"], S["C"]
    fake = g_forward(G, GB, noise, cond)
    loss = -c_forward(C, fake, cond).mean()
    gG = dict(zip(G.keys(), torch.autograd.grad(loss, list(G.values()))))
    if update:
        adamw_step(G, gG, S["adam_g"], hp.lr, hp.betas[0], hp.betas[1], hp.weight_decay)
    return {"generator_loss": loss.item()}, {"G": gG, "fake": fake.detach()}


def synth_batch(hp, B, seed):
    g = torch.Generator().manual_seed(seed)
    real = torch.rand(B, 1, 28, 28, generator=g) * 2 - 1
    labels = torch.randint(hp.num_classes, (B,), generator=g)
    noise = torch.randn(B, hp.latent_size, generator=g)
    alpha = torch.rand(B, 1, generator=g)
    labels2 = torch.randint(hp.num_classes, (B,), generator=g)
    noise2 = torch.randn(B, hp.latent_size, generator=g)
    return {"real": real, "labels": labels, "noise": noise, "alpha": alpha, "labels_g": labels2, "noise_g": noise2}
