"""Oracle (CPU restatement) of MNIST classifier pre-training.  TEST INFRASTRUCTURE.

Follows:
  * ``conditional_counteRGAN/mnist/models/classifier.py:4-28``  (CNNClassifier in TRAIN mode: Dropout2d(0.25) after the
    conv stack, Dropout(0.5) after fc.1)
  * ``conditional_counteRGAN/mnist/trainer.py:8-39``            (train_classifier: Adam(lr=cls_lr), CrossEntropyLoss,
    one optimizer step per batch, validation accuracy per epoch)
  * ``torch/optim/adam.py`` single-tensor Adam (oracle.mnist_countergan.adam_step)
The dropout draws are INPUTS here (keep-masks already scaled by 1/(1-p)): nn.Dropout2d zeroes whole channels per sample
(mask [B,128,1,1]), nn.Dropout single elements (mask [B,256]).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import mnist_countergan as O

P_DROP2D, P_DROP = 0.25, 0.5


def synth_masks(B, seed):
    """Keep-masks as torch's dropout builds them: Bernoulli(1 - p) / (1 - p)."""
    g = torch.Generator().manual_seed(seed)
    m2 = (torch.rand(B, 128, generator=g) >= P_DROP2D).float() / (1.0 - P_DROP2D)
    m1 = (torch.rand(B, 256, generator=g) >= P_DROP).float() / (1.0 - P_DROP)
    return m2, m1


def forward_train(P, x, m2, m1):
    """classifier.py:25-28 in train mode with injected masks."""
    z = F.relu(F.conv2d(x, P["conv.0.weight"], P["conv.0.bias"], stride=1, padding=1))
    z = F.relu(F.conv2d(z, P["conv.2.weight"], P["conv.2.bias"], stride=2, padding=1))
    z = F.relu(F.conv2d(z, P["conv.4.weight"], P["conv.4.bias"], stride=2, padding=1))
    z = z * m2.view(-1, 128, 1, 1)                      # Dropout2d(0.25)
    z = z.flatten(1)
    z = F.relu(z @ P["fc.1.weight"].t() + P["fc.1.bias"])
    z = z * m1                                          # Dropout(0.5)
    return z @ P["fc.4.weight"].t() + P["fc.4.bias"]


def make_state(PC):
    P = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in PC.items())
    return {"C": P, "adam": O.adam_init(P)}


def train_step(S, x, y, m2, m1, lr=1e-3):
    """One iteration of trainer.py:16-21.  Returns (loss, grads)."""
    P = S["C"]
    loss = F.cross_entropy(forward_train(P, x, m2, m1), y)
    grads = torch.autograd.grad(loss, list(P.values()))
    G = OrderedDict(zip(P.keys(), grads))
    O.adam_step(P, G, S["adam"], lr)
    return loss.item(), G


def accuracy(P, loader):
    """trainer.py:24-32."""
    correct = total = 0
    with torch.no_grad():
        for x, y in loader:
            correct += (O.c_forward(P, x).argmax(1) == y).sum().item()
            total += y.size(0)
    return correct / max(total, 1)
