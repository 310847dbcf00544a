"""Oracle (CPU restatement) of the counterfactual evaluation metrics.  TEST INFRASTRUCTURE.
Follows ``conditional_counteRGAN/mnist/eval_utils.py:46-75`` (evaluate_counterfactuals) and ``:78-110``
(evaluate_generator_per_target): generator and classifier in eval mode, mask = ones, x_cf = clamp(x + masked, -1, 1)."""
import torch
import torch.nn.functional as F

from . import mnist_countergan as O


def evaluate_counterfactuals(S, x, y_true, y_target, n_resblocks=6):
    with torch.no_grad():
        _, masked = O.g_forward(S["G"], S["GB"], x, y_target, torch.ones_like(x), n_resblocks=n_resblocks, training=False)
        x_cf = torch.clamp(x + masked, -1.0, 1.0)
        logits = O.c_forward(S["C"], x_cf)
        probs = F.softmax(logits, dim=1)
    idx = torch.arange(len(y_target))
    return {"class_flip_rate": (logits.argmax(1) == y_target).float().mean().item(),
            "prediction_gain": (probs[idx, y_target] - probs[idx, y_true]).mean().item(),
            "actionability": torch.abs(x_cf - x).mean().item()}, x_cf


def per_target(S, batches, num_classes=10, n_resblocks=6):
    res = {c: {"class_flip_rate": [], "prediction_gain": [], "actionability": []} for c in range(num_classes)}
    for x, y in batches:
        for c in range(num_classes):
            m, _ = evaluate_counterfactuals(S, x, y, torch.full_like(y, c), n_resblocks)
            for k in m:
                res[c][k].append(m[k])
    return {c: {k: float(torch.tensor(v).mean()) for k, v in m.items()} for c, m in res.items()}
