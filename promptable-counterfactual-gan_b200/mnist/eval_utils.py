"""Native mirror of the evaluation loops of ``conditional_counteRGAN/mnist/eval_utils.py``:

    evaluate_classifier(classifier, dataloader, device, save_dir=None, prefix)        eval_utils.py:15-43
    evaluate_counterfactuals(generator, classifier, x, y_true, y_target, device)      eval_utils.py:46-75
    evaluate_generator_per_target(generator, classifier, test_loader, config)         eval_utils.py:78-110

Same signatures, return values, prints and CSV file.  The work - for every batch and each of the ten target classes an
eval-mode generator forward (BatchNorm folded into the convolutions, two tcgen05 launches per residual block), the clamp,
the frozen classifier and the three metrics - runs in libpcg: the mirror modules' forwards plus ``pcg_cf_apply`` /
``pcg_cf_metrics``; one device->host read of three floats per (batch, target) instead of the reference's five.
The generator / classifier may be the mirror classes or plain torch modules with the reference's interface.
"""
import os

import torch

from .. import ops as K


def _eval_modules(*mods):
    for m in mods:
        m.eval()


def evaluate_classifier(classifier, dataloader, device, save_dir=None, prefix="classifier"):
    """eval_utils.py:15-43: accuracy + confusion matrix (the heat-map is drawn when matplotlib / seaborn are present)."""
    classifier.eval()
    nc = None
    cm = None
    with torch.no_grad():
        for x, y in dataloader:
            x, y = x.to(device), y.to(device)
            preds = classifier(x).argmax(1)
            if cm is None:
                nc = 10
                cm = torch.zeros(nc, nc, dtype=torch.int64, device=x.device)
            cm += torch.bincount(y * nc + preds, minlength=nc * nc).view(nc, nc)
    cm = cm.cpu()
    acc = cm.diag().sum().item() / max(cm.sum().item(), 1)
    if save_dir:
        os.makedirs(save_dir, exist_ok=True)
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
            import seaborn as sns
        except ImportError:
            plt = None
        if plt is not None:
            plt.figure(figsize=(8, 6))
            sns.heatmap(cm.numpy(), annot=True, fmt="d", cmap="Blues")
            plt.xlabel("Predicted")
            plt.ylabel("True")
            plt.title(f"Confusion Matrix ({prefix}) - Acc={acc:.4f}")
            plt.savefig(os.path.join(save_dir, f"{prefix}_confusion_matrix.png"))
            plt.close()
    print(f"{prefix} Test Accuracy: {acc:.4f}")
    return acc, cm


class _Scratch:
    by_dev = {}

    @classmethod
    def get(cls, dev, shape):
        key = (dev, tuple(shape))
        s = cls.by_dev.get(key)
        if s is None:
            s = cls.by_dev[key] = (torch.empty(shape, device=dev), K.cf_scratch(dev), torch.zeros(3, device=dev))
        return s


def counterfactual_metrics(generator, classifier, x, y_true, y_target):
    """Device-side core of evaluate_counterfactuals: returns (metrics [3] on the device: class-flip rate, prediction
    gain, actionability; x_cf).  x, y_true, y_target are CUDA tensors."""
    _eval_modules(generator, classifier)
    with torch.no_grad():
        x = x.float().contiguous()
        residual = generator(x, y_target, torch.ones_like(x))[1].contiguous()        # masked residual, mask = 1
        x_cf, scratch, out = _Scratch.get(x.device, x.shape)
        K.cf_apply(x, residual, x_cf, scratch, -1.0, 1.0)                            # clamp to the training range
        logits = classifier(x_cf).float().contiguous()
        K.cf_metrics(logits, y_true.long().contiguous(), y_target.long().contiguous(), scratch, x.numel(), out)
    return out, x_cf


def evaluate_counterfactuals(generator, classifier, x, y_true, y_target, device):
    """eval_utils.py:46-75."""
    out, x_cf = counterfactual_metrics(generator, classifier, x.to(device), y_true.to(device), y_target.to(device))
    cfr, gain, act = out.tolist()
    x_vis = ((x + 1.0) / 2.0).detach().cpu()
    x_cf_vis = ((x_cf + 1.0) / 2.0).detach().cpu()
    return {"class_flip_rate": cfr, "prediction_gain": gain, "actionability": act}, (x_vis, x_cf_vis)


def evaluate_generator_per_target(generator, classifier, test_loader, config):
    """eval_utils.py:78-110: per target class, the three metrics averaged over the batches; written to
    ``countergan_metrics_per_class.csv``.  Metrics stay on the device until the end (one read for the whole sweep)."""
    device, num_classes = config.device, config.num_classes
    _eval_modules(generator, classifier)
    sums = torch.zeros(num_classes, 3, device=device)
    nb = 0
    for x, y in test_loader:
        x, y = x.to(device), y.to(device)
        for target_class in range(num_classes):
            out, _ = counterfactual_metrics(generator, classifier, x, y, torch.full_like(y, target_class))
            sums[target_class] += out
        nb += 1
    avg = (sums / max(nb, 1)).cpu()
    results = {cls: {"class_flip_rate": float(avg[cls, 0]), "prediction_gain": float(avg[cls, 1]),
                     "actionability": float(avg[cls, 2])} for cls in range(num_classes)}
    os.makedirs(config.save_dir, exist_ok=True)
    csv_path = os.path.join(config.save_dir, "countergan_metrics_per_class.csv")
    try:
        import pandas as pd
        df = pd.DataFrame.from_dict(results, orient="index")
        df.to_csv(csv_path)
        print(f"Saved per-class CounterGAN metrics to {csv_path}")
        print(df)
    except ImportError:
        with open(csv_path, "w") as f:
            f.write(",class_flip_rate,prediction_gain,actionability\n")
            for cls, m in results.items():
                f.write(f"{cls},{m['class_flip_rate']},{m['prediction_gain']},{m['actionability']}\n")
        print(f"Saved per-class CounterGAN metrics to {csv_path}")
    return results
