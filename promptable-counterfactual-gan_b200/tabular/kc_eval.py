"""Native mirror of the counterfactual evaluation of ``conditional_counteRGAN/house_sales_kc_usa/eval_utils.py``
(SURVEY.md 8f row 2):

    build_counterfactuals(G, x, target_onehot, config) -> (masked_residual, x_cf)          eval_utils.py:25-181
    compute_metrics_per_target(generator, classifier, X, y, config, max_vis=500)           eval_utils.py:185-289

Same signatures and return values.  The generator forward (eval mode: BatchNorm running statistics, FiLM, the seven
categorical heads with HARD Gumbel-softmax samples), the residual assembly (continuous columns, ``one_hot @ norm_vals -
x`` for the categorical ones, immutable-feature mask) and the frozen classifier run as libpcg launches of one ``KcPlan``
of fixed batch size: evaluation batches are ragged (rows whose class equals the target are dropped, eval_utils.py:223),
and because every row is independent in eval mode they are simply padded to the plan's batch.
"""
from collections import OrderedDict

import numpy as np
import torch

from .. import ops as K
from .kc import KcPlan


def _plan(G, classifier, config, batch):
    cache = G.__dict__.setdefault("_pcg_eval_plans", {})
    p = cache.get(batch)
    dev = next(G.parameters()).device
    if p is None or not p.G.aliases(G):
        scaler, norm_vals = config.get('scaler'), None
        if scaler is not None and hasattr(scaler, 'data_min_') and hasattr(scaler, 'data_max_'):
            dmin, dmax = np.array(scaler.data_min_, dtype=float), np.array(scaler.data_max_, dtype=float)
            norm_vals = {int(f): torch.tensor((np.array(info['raw_values'], dtype=float) - dmin[int(f)]) /
                                              (dmax[int(f)] - dmin[int(f)] + 1e-12), dtype=torch.float32)
                         for f, info in config['categorical_info'].items()}          # eval_utils.py:57-66
        p = KcPlan(batch, dev, config['categorical_info'], config['continuous_idx'], norm_vals=norm_vals,
                   input_dim=G.input_dim, hidden=G.hidden_dim, nc=G.num_classes, n_blocks=len(G.blocks), tau=G.tau,
                   use_graph=False)
        p.adopt_g(G)
        cache.clear()
        cache[batch] = p
    if classifier is not None and not p.C.aliases(classifier):
        p.adopt_c(classifier)
    p.refresh()
    return p


def _pad(t, n):
    if t.shape[0] == n:
        return t.contiguous()
    out = torch.zeros((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    out[:t.shape[0]] = t
    return out


def build_counterfactuals(G, x, target_onehot, config, exp_noise=None, classifier=None, _want_plan=False):
    """eval_utils.py:25-181: (masked_residual, x_cf) with x_cf = clamp(x + masked_residual, 0, 1).
    ``exp_noise`` (optional, tests): the Exp(1) draws behind the Gumbel noise, one [bs, n_f] tensor per categorical
    feature; by default they are drawn here, as F.gumbel_softmax does in the reference."""
    bs, d = x.shape
    B = max(int(config.get('batch_size', 128)), bs)
    p = _plan(G, classifier, config, B)
    mask = torch.ones(B, d, device=x.device)
    imm = list(config.get('immutable_idx', []))
    if imm:
        mask[:, imm] = 0.0
    if exp_noise is None:
        exp_noise = [torch.empty(B, n, device=x.device).exponential_() for n in p.cat.values()]
    else:
        exp_noise = [torch.cat([e, torch.ones(B - bs, e.shape[1], device=e.device)]) if e.shape[0] < B else e
                     for e in exp_noise]
    tau = config.get('gumbel_tau', None)
    with torch.no_grad():
        p.g_forward(_pad(x.float(), B), _pad(target_onehot.float(), B), mask, exp_noise,
                    p.tau if tau is None else float(tau), training=False, hard=True)
        p._assemble()                                   # res, masked = res * mask, xcf = x + masked (plan buffers)
        masked = p.masked[:bs].clone()
        x_cf = torch.clamp(x + masked, 0.0, 1.0)
    return (masked, x_cf, p) if _want_plan else (masked, x_cf)


def compute_metrics_per_target(generator, classifier, X, y, config, max_vis=500):
    """eval_utils.py:185-289: per target class the class-flip rate, the prediction gain and the mean |masked residual|
    over the rows of every other class.  Returns (DataFrame, originals, counterfactuals)."""
    device = torch.device(config.get('cuda', 'cuda'))
    if device.type != "cuda":
        raise RuntimeError("pcg_b200.tabular.kc_eval needs a CUDA device (there is no CPU fallback)")
    batch_size = int(config.get('batch_size', 128))
    X_t = torch.tensor(X, dtype=torch.float32, device=device)
    y_t = torch.tensor(y, dtype=torch.long, device=device)
    num_classes = int(np.unique(y).size)
    generator.eval()
    classifier.eval()
    results, originals_vis, cfs_vis = [], [], []
    logit_o = torch.empty(batch_size, num_classes, device=device)
    logit_c = torch.empty(batch_size, num_classes, device=device)
    with torch.no_grad():
        for target in range(num_classes):
            flips, gains, actions = [], [], []
            for i in range(0, X_t.shape[0], batch_size):
                xb, yb = X_t[i:i + batch_size], y_t[i:i + batch_size]
                sel = yb != target
                if int(sel.sum()) == 0:
                    continue
                x = xb[sel]
                bs = x.shape[0]
                t_oh = torch.nn.functional.one_hot(torch.full((bs,), target, device=device), num_classes).float()
                masked, _, p = build_counterfactuals(generator, x, t_oh, config, classifier=classifier, _want_plan=True)
                x_cf = (x + masked)                                          # eval_utils.py:241 (unclamped)
                p._c_fwd(_pad(x, p.B), p.clog0)
                lo = p.clog0[:bs].clone()
                p._c_fwd(_pad(x_cf, p.B), p.clog0)
                lc = p.clog0[:bs]
                po, pc = torch.softmax(lo, 1)[:, target], torch.softmax(lc, 1)[:, target]
                flips.append((lc.argmax(1) == target).float().mean())
                gains.append((pc - po).mean())
                actions.append(masked.abs().mean())
                if len(originals_vis) < max_vis:
                    originals_vis.append(x.cpu())
                    cfs_vis.append(x_cf.cpu())
            agg = (lambda v: float(torch.stack(v).mean()) if v else float('nan'))    # noqa: E731
            results.append({'target_class': int(target), 'class_flip': agg(flips), 'prediction_gain': agg(gains),
                            'avg_actionability': agg(actions)})
    if originals_vis:
        originals_vis, cfs_vis = torch.cat(originals_vis).numpy(), torch.cat(cfs_vis).numpy()
    else:
        originals_vis = cfs_vis = np.empty((0, X.shape[1]))
    import pandas as pd
    return pd.DataFrame(results), originals_vis, cfs_vis
