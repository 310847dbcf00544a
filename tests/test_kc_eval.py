"""KC counterfactual evaluation (house_sales_kc_usa/eval_utils.py:25-289).

CPU : the oracle restatement (oracle.tabular_countergan.kc_build_counterfactuals) against the UNMODIFIED reference
      build_counterfactuals on the reference's own generator (same torch seed => same Gumbel draws);
GPU : the native mirror (pcg_b200/tabular/kc_eval.py) against the oracle: build_counterfactuals with injected noise
      (ragged batch, padded), and compute_metrics_per_target on a generator whose categorical logits dominate the Gumbel
      noise (the hard samples are then deterministic, so the sweep is comparable although it draws its own noise)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import tabular_countergan as T
from tests._refload import experiment


class _Scaler:
    """MinMaxScaler stand-in with the two attributes eval_utils.py:57-60 reads."""

    def __init__(self):
        self.data_min_ = np.zeros(17)
        self.data_max_ = np.ones(17)
        for f, raw in T.KC_RAW.items():
            self.data_min_[f], self.data_max_[f] = min(raw), max(raw)


def _config(batch=64):
    cat = OrderedDict((f, {"n": n, "raw_values": T.KC_RAW[f]}) for f, n in T.KC_CAT.items())
    return {"cuda": "cuda", "batch_size": batch, "categorical_info": cat, "continuous_idx": T.KC_CONT,
            "immutable_idx": T.KC_IMMUTABLE, "scaler": _Scaler(), "gumbel_tau": 0.5}


def _norm_vals(cfg):
    sc = cfg["scaler"]
    return OrderedDict((f, torch.tensor((np.array(info["raw_values"], dtype=float) - sc.data_min_[f]) /
                                        (sc.data_max_[f] - sc.data_min_[f] + 1e-12), dtype=torch.float32))
                       for f, info in cfg["categorical_info"].items())


@pytest.mark.reference
def test_oracle_matches_reference_build_counterfactuals():
    cfg = _config()
    cfg["cuda"] = "cpu"
    x, y, t, mask, noise = T.kc_batch(24, 5)
    oh = torch.nn.functional.one_hot(t, 4).float()
    with experiment("conditional_counteRGAN/house_sales_kc_usa") as imp:
        import sys
        from unittest import mock
        for m in ("seaborn", "eval_utils_mask_analysis", "matplotlib.colors"):      # plotting only; absent here
            sys.modules.setdefault(m, mock.MagicMock())
        ev = imp("eval_utils")
        Gm = imp("models.generator")
        torch.manual_seed(3)
        G = Gm.ResidualGenerator(17, 32, 4, T.KC_CONT, cfg["categorical_info"]).eval()
        with torch.no_grad():
            for n, b in G.named_buffers():
                if n.endswith("running_var"):
                    b.uniform_(0.5, 1.5)
                elif n.endswith("running_mean"):
                    b.normal_(0, 0.2)
        PG = OrderedDict((k, v.detach().clone()) for k, v in G.named_parameters())
        BG = OrderedDict((k, v.detach().clone()) for k, v in G.named_buffers())
        torch.manual_seed(77)
        with torch.no_grad():
            masked_ref, xcf_ref = ev.build_counterfactuals(G, x, oh, cfg)
    torch.manual_seed(77)          # F.gumbel_softmax draws empty_like(logits).exponential_() per head, in ModuleDict order
    noise = [torch.empty(24, n).exponential_() for n in T.KC_CAT.values()]
    masked, xcf = T.kc_build_counterfactuals(PG, BG, x, oh, noise, _norm_vals(cfg))
    assert torch.allclose(masked, masked_ref, atol=1e-6) and torch.allclose(xcf, xcf_ref, atol=1e-6)
    assert torch.all(masked[:, T.KC_IMMUTABLE] == 0)


def _native_modules(scale_heads=1.0):
    import pcg_b200  # noqa: F401
    from pcg_b200.tabular import kc as KC
    cfg = _config()
    torch.manual_seed(3)
    G = KC.ResidualGenerator(17, 32, 4, T.KC_CONT, cfg["categorical_info"])
    C = KC.NNClassifier(17, 4)
    with torch.no_grad():
        for m in (G, C):
            for n, b in m.named_buffers():
                if n.endswith("running_var"):
                    b.uniform_(0.5, 1.5)
                elif n.endswith("running_mean"):
                    b.normal_(0, 0.2)
        for n, p in G.named_parameters():
            if n.startswith("fc_cat_logits") and n.endswith("weight"):
                p.mul_(scale_heads)
    PG = OrderedDict((k, v.detach().clone()) for k, v in G.named_parameters())
    BG = OrderedDict((k, v.detach().clone()) for k, v in G.named_buffers())
    PC = OrderedDict((k, v.detach().clone()) for k, v in C.named_parameters())
    BC = OrderedDict((k, v.detach().clone()) for k, v in C.named_buffers())
    return cfg, G.cuda().eval(), C.cuda().eval(), PG, BG, PC, BC


@pytest.mark.gpu
def test_native_build_counterfactuals_matches_oracle():
    from pcg_b200.tabular import kc_eval as EV
    cfg, G, C, PG, BG, PC, BC = _native_modules()
    x, y, t, mask, noise = T.kc_batch(37, 5)                         # ragged: padded to the plan's batch of 64
    oh = torch.nn.functional.one_hot(t, 4).float()
    masked, xcf = EV.build_counterfactuals(G, x.cuda(), oh.cuda(), cfg, exp_noise=[e.cuda() for e in noise])
    want_m, want_x = T.kc_build_counterfactuals(PG, BG, x, oh, noise, _norm_vals(cfg))
    assert masked.shape == (37, 17)
    assert (masked.cpu() - want_m).abs().max() < 2e-5 and (xcf.cpu() - want_x).abs().max() < 2e-5
    assert torch.all(masked[:, T.KC_IMMUTABLE] == 0) and xcf.min() >= 0 and xcf.max() <= 1
    # hard=True through the module's own forward
    cont, logits, samples = G(x.cuda(), oh.cuda(), hard=True)
    for f, s in samples.items():
        assert torch.all((s == 0) | (s == 1)) and torch.all(s.sum(1) == 1)


@pytest.mark.gpu
def test_native_metrics_per_target_match_oracle():
    from pcg_b200.tabular import kc_eval as EV
    cfg, G, C, PG, BG, PC, BC = _native_modules(scale_heads=3000.0)   # logits >> Gumbel noise: deterministic hard samples
    X = torch.cat([T.kc_batch(64, 200 + i)[0] for i in range(3)])
    y = torch.randint(0, 4, (192,), generator=torch.Generator().manual_seed(1))
    df, orig, cfs = EV.compute_metrics_per_target(G, C, X.numpy(), y.numpy(), cfg, max_vis=5)
    ones = lambda bs: [torch.ones(bs, n) for n in T.KC_CAT.values()]       # noqa: E731  (noise is irrelevant here)
    want = T.kc_metrics_per_target(PG, BG, PC, BC, X, y, _norm_vals(cfg), ones, batch_size=64)
    assert list(df.columns) == ["target_class", "class_flip", "prediction_gain", "avg_actionability"] and len(df) == 4
    for row, w in zip(df.to_dict("records"), want):
        assert abs(row["class_flip"] - w["class_flip"]) <= 0.02, (row, w)
        # the scaled head logits also amplify fp32 rounding: a near-tie between two categories flips in a few rows
        assert abs(row["prediction_gain"] - w["prediction_gain"]) <= 1e-3, (row, w)
        assert abs(row["avg_actionability"] - w["avg_actionability"]) <= 1e-3, (row, w)
    assert orig.shape[1] == 17 and orig.shape == cfs.shape and orig.shape[0] >= 5
