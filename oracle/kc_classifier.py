"""Oracle (CPU restatement) of the KC house-sales classifier pre-training iteration.  TEST INFRASTRUCTURE.

Follows:
  * ``conditional_counteRGAN/house_sales_kc_usa/models/nn_classifier.py:4-32`` (NNClassifier in TRAIN mode:
    Linear -> LeakyReLU(0.1) -> BatchNorm1d (batch statistics, running buffers updated) -> Dropout(0.3 / 0.2 / 0.1 / none))
  * ``conditional_counteRGAN/house_sales_kc_usa/trainer.py:56-61,85-96`` (CrossEntropyLoss(weight=class_weights),
    AdamW(lr, weight_decay), one optimizer step per batch, running loss / accuracy counters)
  * ``torch/optim/adamw.py`` (decoupled decay: p *= 1 - lr*wd, then the Adam update of ``oracle.mnist_countergan.adam_step``)
Dropout keep-masks (scaled by 1/(1-p)) are inputs.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import mnist_countergan as O

DIMS = [256, 256, 128, 64]
P_DROP = [0.3, 0.2, 0.1, 0.0]
LIN = ["net.0", "net.4", "net.8", "net.12", "net.15"]
BN = ["net.2", "net.6", "net.10", "net.14"]


def shapes(input_dim=17, out_dim=4):
    s, a = OrderedDict(), input_dim
    for j, b in enumerate(DIMS):
        s[LIN[j] + ".weight"], s[LIN[j] + ".bias"] = (b, a), (b,)
        s[BN[j] + ".weight"], s[BN[j] + ".bias"] = (b,), (b,)
        a = b
    s[LIN[4] + ".weight"], s[LIN[4] + ".bias"] = (out_dim, a), (out_dim,)
    return s


def buffers():
    b = OrderedDict()
    for j, c in enumerate(DIMS):
        b[BN[j] + ".running_mean"], b[BN[j] + ".running_var"] = torch.zeros(c), torch.ones(c)
        b[BN[j] + ".num_batches_tracked"] = torch.tensor(0)
    return b


def synth_masks(B, seed):
    g = torch.Generator().manual_seed(seed)
    return [(torch.rand(B, c, generator=g) >= p).float() / (1.0 - p) for c, p in zip(DIMS[:3], P_DROP[:3])]


def forward(P, Bf, x, masks=None, training=True):
    h = x
    for j in range(4):
        h = F.leaky_relu(h @ P[LIN[j] + ".weight"].t() + P[LIN[j] + ".bias"], 0.1)
        if training:
            mean, var = h.mean(0), h.var(0, unbiased=False)
            with torch.no_grad():
                n = h.shape[0]
                Bf[BN[j] + ".running_mean"].mul_(0.9).add_(0.1 * mean)
                Bf[BN[j] + ".running_var"].mul_(0.9).add_(0.1 * var * (n / max(n - 1, 1)))
                Bf[BN[j] + ".num_batches_tracked"] += 1
        else:
            mean, var = Bf[BN[j] + ".running_mean"], Bf[BN[j] + ".running_var"]
        h = (h - mean) / torch.sqrt(var + 1e-5) * P[BN[j] + ".weight"] + P[BN[j] + ".bias"]
        if training and j < 3 and masks is not None:
            h = h * masks[j]
    return h @ P[LIN[4] + ".weight"].t() + P[LIN[4] + ".bias"]


def make_state(PC):
    P = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in PC.items())
    return {"C": P, "B": buffers(), "adam": O.adam_init(P)}


def train_step(S, x, y, masks, class_weights, lr=1e-3, wd=1e-4):
    """trainer.py:85-96.  Returns (loss, correct predictions, grads)."""
    P = S["C"]
    logits = forward(P, S["B"], x, masks, True)
    loss = F.cross_entropy(logits, y, weight=class_weights)
    grads = torch.autograd.grad(loss, list(P.values()))
    G = OrderedDict(zip(P.keys(), grads))
    with torch.no_grad():
        for p in P.values():
            p.mul_(1.0 - lr * wd)                              # AdamW: decoupled weight decay
    O.adam_step(P, G, S["adam"], lr)
    return loss.item(), int((logits.argmax(1) == y).sum()), G
