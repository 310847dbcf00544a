"""Evaluation loops (conditional_counteRGAN/mnist/eval_utils.py:15-110).

CPU : the oracle (oracle/mnist_eval.py) against the UNMODIFIED reference evaluate_counterfactuals /
      evaluate_generator_per_target run here on the reference's own modules;
GPU : the native mirror (pcg_b200/mnist/eval_utils.py: eval-mode generator with folded BatchNorm, pcg_cf_apply /
      pcg_cf_metrics) against the oracle on the same weights, in fp32 (tight) and in the default bf16 mode."""
import types
from collections import OrderedDict

import pytest
import torch

from oracle import mnist_countergan as O
from oracle import mnist_eval as OE
from tests._refload import experiment


def _state_from(G, C, D=None):
    return {"G": OrderedDict((k, v.detach().cpu()) for k, v in G.named_parameters()),
            "GB": OrderedDict((k, v.detach().cpu().clone()) for k, v in G.named_buffers()),
            "C": OrderedDict((k, v.detach().cpu()) for k, v in C.named_parameters())}


def _randomize(G):
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for n, b in G.named_buffers():
            if n.endswith("running_mean"):
                b.copy_(0.2 * torch.randn(b.shape, generator=g))
            elif n.endswith("running_var"):
                b.copy_(0.5 + torch.rand(b.shape, generator=g))
        for n, p in G.named_parameters():
            if ".bn" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            if n == "conv_out.weight":
                p.mul_(6.0)                      # residuals large enough to move the classifier


@pytest.mark.reference
def test_oracle_matches_reference_eval_functions(tmp_path):
    batches = [O.synth_batch(6, 300 + i, mnist_like=True)[:2] for i in range(2)]
    with experiment("conditional_counteRGAN/mnist") as imp:
        import sys
        from unittest import mock
        for m in ("seaborn", "tqdm", "sklearn.metrics"):
            sys.modules.setdefault(m, mock.MagicMock())
        cfgm = types.ModuleType("config")
        cfgm.Config = type("Config", (), dict(num_classes=10))
        sys.modules["config"] = cfgm
        ev = imp("eval_utils")
        ev.tqdm = lambda it, **k: it
        Gm, Cm = imp("models.generator"), imp("models.classifier")
        torch.manual_seed(4)
        G, C = Gm.ResidualGenerator(base_ch=16, n_resblocks=2), Cm.CNNClassifier()
        _randomize(G)
        S = _state_from(G, C)
        x, y = batches[0]
        t = torch.randint(0, 10, (6,), generator=torch.Generator().manual_seed(0))
        ref, (xv, xcv) = ev.evaluate_counterfactuals(G, C, x, y, t, "cpu")
        got, x_cf = OE.evaluate_counterfactuals(S, x, y, t, n_resblocks=2)
        for k in ref:
            assert abs(got[k] - ref[k]) < 1e-6, (k, got[k], ref[k])
        assert torch.allclose((x_cf + 1) / 2, xcv, atol=1e-6)
        cfg = types.SimpleNamespace(device="cpu", num_classes=10, save_dir=str(tmp_path))
        ev.evaluate_generator_per_target(G, C, batches, cfg)
        import pandas as pd
        df = pd.read_csv(tmp_path / "countergan_metrics_per_class.csv", index_col=0)
        want = OE.per_target(S, batches, n_resblocks=2)
        for c in range(10):
            for k in want[c]:
                assert abs(df.loc[c, k] - want[c][k]) < 1e-6, (c, k)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,ch,tol", [("fp32", 16, 2e-4), ("bf16", 64, 3e-2)])
def test_native_eval_loops_match_oracle(tmp_path, precision, ch, tol):
    import pcg_b200  # noqa: F401
    from pcg_b200.mnist import eval_utils as EV
    from pcg_b200.mnist.models.classifier import CNNClassifier
    from pcg_b200.mnist.models.generator import ResidualGenerator
    torch.manual_seed(4)
    G, C = ResidualGenerator(base_ch=ch, n_resblocks=2), CNNClassifier()
    _randomize(G)
    S = _state_from(G, C)
    G, C = G.cuda(), C.cuda()
    G.precision = C.precision = precision
    batches = [O.synth_batch(16, 300 + i, mnist_like=True)[:2] for i in range(2)]
    x, y = batches[0]
    t = torch.randint(0, 10, (16,), generator=torch.Generator().manual_seed(0))
    got, (xv, xcv) = EV.evaluate_counterfactuals(G, C, x, y, t, "cuda")
    want, x_cf = OE.evaluate_counterfactuals(S, x, y, t, n_resblocks=2)
    assert abs(got["actionability"] - want["actionability"]) <= tol * want["actionability"] + 1e-6
    assert abs(got["prediction_gain"] - want["prediction_gain"]) <= tol * 5 + 1e-6
    assert abs(got["class_flip_rate"] - want["class_flip_rate"]) <= (0.0 if precision == "fp32" else 2 / 16) + 1e-6
    assert ((xcv - (x_cf + 1) / 2).abs().max() / 1.0).item() < tol
    cfg = types.SimpleNamespace(device="cuda", num_classes=10, save_dir=str(tmp_path))
    res = EV.evaluate_generator_per_target(G, C, batches, cfg)
    ref = OE.per_target(S, batches, n_resblocks=2)
    for c in range(10):
        assert abs(res[c]["actionability"] - ref[c]["actionability"]) <= tol * ref[c]["actionability"] + 1e-6, c
        assert abs(res[c]["prediction_gain"] - ref[c]["prediction_gain"]) <= tol * 5 + 1e-6, c
    assert (tmp_path / "countergan_metrics_per_class.csv").exists()
    acc, cm = EV.evaluate_classifier(C, batches, "cuda")
    with torch.no_grad():
        want_acc = sum(int((O.c_forward(S["C"], xb).argmax(1) == yb).sum()) for xb, yb in batches) / 32
    assert abs(acc - want_acc) <= (0.0 if precision == "fp32" else 1 / 32) + 1e-9 and int(cm.sum()) == 32
