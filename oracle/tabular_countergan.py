"""Oracle (CPU restatement) of the two tabular CounteRGAN iterations.  TEST INFRASTRUCTURE.

  moons:  conditional_counteRGAN/moons/trainer.py:58-96, models/{generator,discriminator,nn_classifier}.py
  KC:     conditional_counteRGAN/house_sales_kc_usa/trainer.py:241-316, models/generator.py:6-92 (FiLM, ResidualBlock,
          ResidualGenerator with Gumbel-softmax heads), models/discriminator.py:5-20, models/nn_classifier.py:4-32
  spectral norm: torch/nn/utils/spectral_norm.py:62-113 (one power iteration per forward in train mode, in-place u/v,
          sigma = u^T W v, weight = weight_orig / sigma, u and v constants for the gradient)

All random draws are injected: target classes, feature masks, and (KC) the Exp(1) samples torch's
``F.gumbel_softmax`` turns into Gumbel noise.  State dicts use the reference's ``state_dict`` key names.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

from .mnist_countergan import adam_init, adam_step, batch_norm_eval, cross_entropy

KC_CAT = OrderedDict([(0, 9), (1, 30), (4, 6), (5, 2), (6, 5), (7, 5), (8, 13)])        # config.py:57-79
KC_RAW = {0: list(range(9)),
          1: sorted([0.00, 0.50, 0.75, 1.00, 1.25, 1.50, 1.75, 2.00, 2.25, 2.50, 2.75, 3.00, 3.25, 3.50, 3.75, 4.00, 4.25,
                     4.50, 4.75, 5.00, 5.25, 5.50, 5.75, 6.00, 6.25, 6.50, 6.75, 7.50, 7.75, 8.00]),
          4: [1.0, 1.5, 2.0, 2.5, 3.0, 3.5], 5: [0, 1], 6: [0, 1, 2, 3, 4], 7: [1, 2, 3, 4, 5], 8: list(range(1, 14))}
KC_CONT = [i for i in range(17) if i not in KC_CAT]                                      # config.py:82
KC_IMMUTABLE = [13, 14, 11, 12]                                                          # lat, long, yr_built, yr_renovated


# --------------------------------------------------------------------------- shapes / synthetic parameters
def moons_shapes(input_dim=2, hidden=32, nc=3):
    g = OrderedDict()
    dims = [(2 * input_dim + nc, hidden), (hidden, hidden), (hidden, hidden // 2)]
    for i, (a, b) in enumerate(dims):
        g[f"net.{3 * i}.weight"], g[f"net.{3 * i}.bias"] = (b, a), (b,)
        g[f"net.{3 * i + 1}.weight"], g[f"net.{3 * i + 1}.bias"] = (b,), (b,)
    g["net.9.weight"], g["net.9.bias"] = (input_dim, hidden // 2), (input_dim,)
    d = sn_shapes([(input_dim + nc, hidden), (hidden, hidden // 2), (hidden // 2, hidden // 2), (hidden // 2, 1)])
    c = OrderedDict([("net.0.weight", (hidden, input_dim)), ("net.0.bias", (hidden,)), ("net.2.weight", (hidden, hidden)),
                     ("net.2.bias", (hidden,)), ("net.4.weight", (nc, hidden)), ("net.4.bias", (nc,))])
    return g, d, c


def sn_shapes(dims):
    """spectral_norm(nn.Linear): parameters (bias, weight_orig) — bias first, weight_orig is registered later."""
    d = OrderedDict()
    for i, (a, b) in enumerate(dims):
        d[f"net.{2 * i}.bias"], d[f"net.{2 * i}.weight_orig"] = (b,), (b, a)
    return d


def sn_buffers(dims, seed):
    g = torch.Generator().manual_seed(seed)
    b = OrderedDict()
    for i, (a, bb) in enumerate(dims):
        b[f"net.{2 * i}.weight_u"] = F.normalize(torch.randn(bb, generator=g), dim=0, eps=1e-12)
        b[f"net.{2 * i}.weight_v"] = F.normalize(torch.randn(a, generator=g), dim=0, eps=1e-12)
    return b


def kc_shapes(input_dim=17, hidden=32, nc=4, n_blocks=5):
    cond = input_dim + nc
    g = OrderedDict([("fc_in.weight", (hidden, input_dim + cond)), ("fc_in.bias", (hidden,))])
    for i in range(n_blocks):
        p = f"blocks.{i}."
        for nm, shp in (("fc1.weight", (hidden, hidden)), ("fc1.bias", (hidden,)), ("bn1.weight", (hidden,)),
                        ("bn1.bias", (hidden,)), ("fc2.weight", (hidden, hidden)), ("fc2.bias", (hidden,)),
                        ("bn2.weight", (hidden,)), ("bn2.bias", (hidden,)), ("film.gamma.weight", (hidden, cond)),
                        ("film.gamma.bias", (hidden,)), ("film.beta.weight", (hidden, cond)), ("film.beta.bias", (hidden,))):
            g[p + nm] = shp
    g["fc_cont.weight"], g["fc_cont.bias"] = (len(KC_CONT), hidden), (len(KC_CONT),)
    for idx, n in KC_CAT.items():
        g[f"fc_cat_logits.{idx}.weight"], g[f"fc_cat_logits.{idx}.bias"] = (n, hidden), (n,)
    d = sn_shapes([(input_dim + nc, hidden), (hidden, 2 * hidden), (2 * hidden, 4 * hidden), (4 * hidden, 1)])
    c = OrderedDict()
    dims = [(input_dim, 256), (256, 256), (256, 128), (128, 64)]
    idx = 0
    for j, (a, b) in enumerate(dims):
        c[f"net.{idx}.weight"], c[f"net.{idx}.bias"] = (b, a), (b,)
        c[f"net.{idx + 2}.weight"], c[f"net.{idx + 2}.bias"] = (b,), (b,)        # BatchNorm1d after the LeakyReLU
        idx += 4 if j < 3 else 3
    c[f"net.{idx}.weight"], c[f"net.{idx}.bias"] = (nc, 64), (nc,)
    return g, d, c


def kc_d_dims(input_dim=17, hidden=32, nc=4):
    return [(input_dim + nc, hidden), (hidden, 2 * hidden), (2 * hidden, 4 * hidden), (4 * hidden, 1)]


def moons_d_dims(input_dim=2, hidden=32, nc=3):
    return [(input_dim + nc, hidden), (hidden, hidden // 2), (hidden // 2, hidden // 2), (hidden // 2, 1)]


def synth_params(shapes, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for k, s in shapes.items():
        if len(s) == 2:
            b = 1.0 / math.sqrt(s[1])
            a = rng.uniform(-b, b, size=s)
        elif ("bn" in k or _is_bn_key(k, shapes)) and k.endswith("weight"):
            a = 1.0 + 0.1 * rng.standard_normal(s)
        else:
            a = 0.1 * rng.standard_normal(s)
        out[k] = torch.from_numpy(a.astype("float32"))
    return out


def _is_bn_key(k, shapes):
    return k.endswith("weight") and len(shapes[k]) == 1


def bn_buffers(shapes, seed=None, randomize=False):
    """running_mean / running_var / num_batches_tracked for every 1-D '*.weight' (BatchNorm) entry."""
    g = torch.Generator().manual_seed(seed or 0)
    b = OrderedDict()
    for k, s in shapes.items():
        if k.endswith(".weight") and len(s) == 1:
            n = k[:-7]
            b[n + ".running_mean"] = torch.randn(s, generator=g) * 0.2 if randomize else torch.zeros(s)
            b[n + ".running_var"] = torch.rand(s, generator=g) + 0.5 if randomize else torch.ones(s)
            b[n + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return b


# --------------------------------------------------------------------------- shared ops
def bn1d_train(x, gamma, beta, rm, rv, nbt, eps=1e-5, mom=0.1):
    n = x.shape[0]
    mean = x.mean(0)
    var = ((x - mean) ** 2).mean(0)
    with torch.no_grad():
        rm.mul_(1 - mom).add_(mom * mean.detach())
        rv.mul_(1 - mom).add_(mom * var.detach() * (n / max(n - 1, 1)))
        nbt.add_(1)
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def sn_weight(P, Bf, name, iterate=True, eps=1e-12):
    """torch.nn.utils.spectral_norm.compute_weight with n_power_iterations = 1."""
    W = P[name + ".weight_orig"]
    u, v = Bf[name + ".weight_u"], Bf[name + ".weight_v"]
    if iterate:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(W.t(), u), dim=0, eps=eps))
            u.copy_(F.normalize(torch.mv(W, v), dim=0, eps=eps))
    uu, vv = u.clone(), v.clone()
    sigma = torch.dot(uu, torch.mv(W, vv))
    return W / sigma


def d_forward(P, Bf, x, onehot, iterate=True):
    """Spectral-norm MLP critic (moons/models/discriminator.py:19-22, house_sales_kc_usa/models/discriminator.py:18-20)."""
    h = torch.cat([x, onehot], 1)
    for i in range(4):
        n = f"net.{2 * i}"
        h = h @ sn_weight(P, Bf, n, iterate).t() + P[n + ".bias"]
        if i < 3:
            h = F.leaky_relu(h, 0.2)
    return h


# --------------------------------------------------------------------------- moons
def moons_g_forward(P, Bf, x, onehot, mask, training=True):
    h = torch.cat([x, onehot, mask], 1)
    for i in range(3):
        h = h @ P[f"net.{3 * i}.weight"].t() + P[f"net.{3 * i}.bias"]
        q = f"net.{3 * i + 1}"
        if training:
            h = bn1d_train(h, P[q + ".weight"], P[q + ".bias"], Bf[q + ".running_mean"], Bf[q + ".running_var"],
                           Bf[q + ".num_batches_tracked"])
        else:
            h = (h - Bf[q + ".running_mean"]) * torch.rsqrt(Bf[q + ".running_var"] + 1e-5) * P[q + ".weight"] + P[q + ".bias"]
        h = torch.relu(h)
    raw = h @ P["net.9.weight"].t() + P["net.9.bias"]
    return raw, raw * mask


def moons_c_forward(P, x):
    h = torch.relu(x @ P["net.0.weight"].t() + P["net.0.bias"])
    h = torch.relu(h @ P["net.2.weight"].t() + P["net.2.bias"])
    return h @ P["net.4.weight"].t() + P["net.4.bias"]


def make_state(PG, BG, PD, BD, PC, BC=None):
    def cp(d, grad):
        o = OrderedDict()
        for k, v in (d or {}).items():
            t = v.detach().clone()
            if t.is_floating_point():
                t.requires_grad_(grad)
            o[k] = t
        return o
    S = {"G": cp(PG, True), "GB": cp(BG, False), "D": cp(PD, True), "DB": cp(BD, False), "C": cp(PC, False),
         "CB": cp(BC, False)}
    S["adam_g"], S["adam_d"] = adam_init(S["G"]), adam_init(S["D"])
    return S


def moons_step(S, x, y, target, mask, nc=3, lr_g=1e-3, lr_d=1e-3, lam=(2.0, 5.0, 5.0, 3.0)):
    """moons/trainer.py:58-96 (target :59-61 and mask :64 injected)."""
    G, GB, D, DB, C = S["G"], S["GB"], S["D"], S["DB"], S["C"]
    t_oh = F.one_hot(target, nc).float()
    raw, masked = moons_g_forward(G, GB, x, t_oh, mask)
    mask_pen = torch.mean(torch.abs(raw * (1.0 - mask)))
    x_cf = x + masked
    D_real = d_forward(D, DB, x, F.one_hot(y, nc).float())
    D_fake = d_forward(D, DB, x_cf.detach(), t_oh)
    D_loss = -D_real.mean() + D_fake.mean()
    gD = dict(zip(D.keys(), torch.autograd.grad(D_loss, list(D.values()))))
    adam_step(D, gD, S["adam_d"], lr_d)
    D_g = d_forward(D, DB, x_cf, t_oh)
    adv = -D_g.mean()
    cls = cross_entropy(moons_c_forward(C, x_cf), target)
    l1 = torch.mean(torch.norm(masked, p=1, dim=1))
    l2 = torch.mean(torch.norm(masked, p=2, dim=1))
    G_loss = adv + lam[0] * cls + lam[1] * l1 + lam[2] * l2 + lam[3] * mask_pen
    gG = dict(zip(G.keys(), torch.autograd.grad(G_loss, list(G.values()))))
    adam_step(G, gG, S["adam_g"], lr_g)
    sc = {"d_loss": D_loss.item(), "g_loss": G_loss.item(), "g_adv": adv.item(), "g_cls": cls.item(), "l1": l1.item(),
          "l2": l2.item(), "mask_pen": mask_pen.item()}
    return sc, {"D": gD, "G": gG, "x_cf": x_cf.detach(), "raw": raw.detach()}


def moons_batch(B, seed, nc=3):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 2, generator=g)
    y = torch.randint(0, nc, (B,), generator=g)
    t = torch.randint(0, nc, (B,), generator=g)
    t = torch.where(t == y, (t + 1) % nc, t)
    mask = torch.randint(0, 2, (B, 2), generator=g).float()
    return x, y, t, mask


# --------------------------------------------------------------------------- KC house sales
def kc_norm_vals(scaler_min=None, scaler_range=None):
    """cat_norm_maps of trainer.py:205-224; without a scaler the /(n-1) fallback (:217-224)."""
    out = OrderedDict()
    for f, n in KC_CAT.items():
        if scaler_min is None:
            out[f] = torch.arange(n, dtype=torch.float32) / max(1.0, n - 1)
        else:
            raw = torch.tensor(KC_RAW[f], dtype=torch.float64)
            out[f] = ((raw - scaler_min[f]) / (scaler_range[f] + 1e-12)).float()
    return out


def kc_g_forward(P, Bf, x, onehot, mask, exp_noise, tau=0.5, training=True):
    """ResidualGenerator.forward (generator.py:68-92).  exp_noise: list of Exp(1) tensors in ModuleDict order."""
    cond = torch.cat([onehot, mask], 1)
    h = torch.relu(torch.cat([x, cond], 1) @ P["fc_in.weight"].t() + P["fc_in.bias"])
    n_blocks = len([k for k in P if k.endswith("fc1.weight")])
    for i in range(n_blocks):
        p = f"blocks.{i}."
        g = cond @ P[p + "film.gamma.weight"].t() + P[p + "film.gamma.bias"]
        b = cond @ P[p + "film.beta.weight"].t() + P[p + "film.beta.bias"]

        def bn(t, j):
            q = p + f"bn{j}"
            if training:
                return bn1d_train(t, P[q + ".weight"], P[q + ".bias"], Bf[q + ".running_mean"], Bf[q + ".running_var"],
                                  Bf[q + ".num_batches_tracked"])
            return (t - Bf[q + ".running_mean"]) * torch.rsqrt(Bf[q + ".running_var"] + 1e-5) * P[q + ".weight"] + P[q + ".bias"]
        out = bn(h @ P[p + "fc1.weight"].t() + P[p + "fc1.bias"], 1)
        out = torch.relu(g * out + b)
        out = bn(out @ P[p + "fc2.weight"].t() + P[p + "fc2.bias"], 2)
        out = g * out + b
        h = h + out
    cont = (h @ P["fc_cont.weight"].t() + P["fc_cont.bias"]) * 0.1
    samples = OrderedDict()
    for (f, n), e in zip(KC_CAT.items(), exp_noise):
        logits = h @ P[f"fc_cat_logits.{f}.weight"].t() + P[f"fc_cat_logits.{f}.bias"]
        samples[f] = torch.softmax((logits - torch.log(e)) / tau, dim=-1)      # F.gumbel_softmax, hard=False
    return cont, samples


def kc_c_forward(P, Bf, x):
    """NNClassifier in eval mode: Linear -> LeakyReLU(0.1) -> BatchNorm1d(running stats) -> (Dropout = id)."""
    h = x
    idx = 0
    for j in range(4):
        h = F.leaky_relu(h @ P[f"net.{idx}.weight"].t() + P[f"net.{idx}.bias"], 0.1)
        q = f"net.{idx + 2}"
        h = (h - Bf[q + ".running_mean"]) * torch.rsqrt(Bf[q + ".running_var"] + 1e-5) * P[q + ".weight"] + P[q + ".bias"]
        idx += 4 if j < 3 else 3
    return h @ P[f"net.{idx}.weight"].t() + P[f"net.{idx}.bias"]


def kc_assemble(x, cont, samples, norm_vals):
    """residual_full of trainer.py:266-282."""
    res = torch.zeros_like(x)
    cols = [None] * x.shape[1]
    for i, f in enumerate(KC_CONT):
        cols[f] = cont[:, i]
    for f, s in samples.items():
        cols[f] = s.matmul(norm_vals[f]) - x[:, f]
    return torch.stack(cols, 1) + 0 * res


def kc_step(S, x, y, target, mask, exp_noise, norm_vals, nc=4, lr_g=1e-3, lr_d=1e-3, lam=(2.0, 1.0, 1.0), tau=0.5):
    """house_sales_kc_usa/trainer.py:241-316 with the draws of :248-254 and the Gumbel noise injected."""
    G, GB, D, DB, C, CB = S["G"], S["GB"], S["D"], S["DB"], S["C"], S["CB"]
    t_oh = F.one_hot(target, nc).float()
    cont, samples = kc_g_forward(G, GB, x, t_oh, mask, exp_noise, tau)
    res = kc_assemble(x, cont, samples, norm_vals)
    masked = res * mask
    x_cf = x + masked
    mask_pen = torch.mean(torch.abs(res * (1.0 - mask)))
    D_real = d_forward(D, DB, x, F.one_hot(y, nc).float())
    D_fake = d_forward(D, DB, x_cf.detach(), t_oh)
    D_loss = -D_real.mean() + D_fake.mean()
    gD = dict(zip(D.keys(), torch.autograd.grad(D_loss, list(D.values()))))
    adam_step(D, gD, S["adam_d"], lr_d)
    D_g = d_forward(D, DB, x_cf, t_oh)
    adv = -D_g.mean()
    logits_cf = kc_c_forward(C, CB, x_cf)
    cls = cross_entropy(logits_cf, target)
    reg = torch.mean(torch.norm(masked, p=1, dim=1))
    G_loss = adv + lam[0] * cls + lam[1] * reg + lam[2] * mask_pen
    gG = dict(zip(G.keys(), torch.autograd.grad(G_loss, list(G.values()), allow_unused=True)))
    adam_step(G, gG, S["adam_g"], lr_g)
    with torch.no_grad():       # diagnostics of trainer.py:319-343
        p_o = torch.softmax(kc_c_forward(C, CB, x), 1)[torch.arange(x.shape[0]), target]
        p_c = torch.softmax(logits_cf, 1)[torch.arange(x.shape[0]), target]
        diag = {"pred_gain": (p_c - p_o).mean().item(),
                "sparsity": 1.0 - (masked.abs() > 1e-3).float().mean().item(),
                "l2": torch.mean(torch.norm(masked, p=2, dim=1)).item(),
                "flip": (logits_cf.argmax(1) == target).float().mean().item()}
    sc = {"d_loss": D_loss.item(), "g_loss": G_loss.item(), "g_adv": adv.item(), "g_cls": cls.item(), "reg": reg.item(),
          "mask_pen": mask_pen.item(), **diag}
    return sc, {"D": gD, "G": gG, "x_cf": x_cf.detach(), "res": res.detach()}


def kc_batch(B, seed, nc=4):
    """x ~ U(0,1) with categorical columns snapped to their normalised grid, immutable columns masked."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 17, generator=g)
    nv = kc_norm_vals()
    for f, n in KC_CAT.items():
        x[:, f] = nv[f][torch.randint(0, n, (B,), generator=g)]
    y = torch.randint(0, nc, (B,), generator=g)
    t = torch.randint(0, nc, (B,), generator=g)
    t = torch.where(t == y, (t + 1) % nc, t)
    mask = torch.randint(0, 2, (B, 17), generator=g).float()
    mask[:, KC_IMMUTABLE] = 0.0
    noise = [torch.empty(B, n).exponential_(generator=g) for n in KC_CAT.values()]
    return x, y, t, mask, noise


# ---------------------------------------------------------------------------------------------- evaluation
def kc_build_counterfactuals(P, Bf, x, onehot, exp_noise, norm_vals, immutable_idx=KC_IMMUTABLE, tau=0.5):
    """house_sales_kc_usa/eval_utils.py:25-181: generator in eval mode with HARD Gumbel-softmax samples, residual
    assembly, immutable-feature mask; returns (masked_residual, x_cf = clamp(x + masked, 0, 1))."""
    mask = torch.ones_like(x)
    if len(immutable_idx):
        mask[:, list(immutable_idx)] = 0.0
    with torch.no_grad():
        cont, soft = kc_g_forward(P, Bf, x, onehot, mask, exp_noise, tau=tau, training=False)
        hard = OrderedDict((f, F.one_hot(s.argmax(1), s.shape[1]).float()) for f, s in soft.items())
        res = kc_assemble(x, cont, hard, norm_vals)
        masked = res * mask
    return masked, torch.clamp(x + masked, 0.0, 1.0)


def kc_metrics_per_target(PG, BG, PC, BC, X, y, norm_vals, noise_fn, batch_size=128, nc=4):
    """eval_utils.py:185-289 (the three per-target metrics); noise_fn(bs) -> list of Exp(1) tensors."""
    out = []
    with torch.no_grad():
        for target in range(nc):
            flips, gains, acts = [], [], []
            for i in range(0, X.shape[0], batch_size):
                xb, yb = X[i:i + batch_size], y[i:i + batch_size]
                sel = yb != target
                if int(sel.sum()) == 0:
                    continue
                x = xb[sel]
                bs = x.shape[0]
                oh = F.one_hot(torch.full((bs,), target), nc).float()
                masked, _ = kc_build_counterfactuals(PG, BG, x, oh, noise_fn(bs), norm_vals)
                x_cf = x + masked
                po = torch.softmax(kc_c_forward(PC, BC, x), 1)[:, target]
                lc = kc_c_forward(PC, BC, x_cf)
                pc = torch.softmax(lc, 1)[:, target]
                flips.append((lc.argmax(1) == target).float().mean().item())
                gains.append((pc - po).mean().item())
                acts.append(masked.abs().mean().item())
            out.append({"target_class": target, "class_flip": sum(flips) / len(flips), "prediction_gain": sum(gains) / len(gains),
                        "avg_actionability": sum(acts) / len(acts)})
    return out
