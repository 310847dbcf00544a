"""Golden vectors of the two tabular CounteRGANs produced by the reference's own ``train_countergan`` functions
(tests/golden/make_golden_tabular.py; moons: 3 iterations, KC house sales: 2 iterations, batch 16):
  * CPU (-m "not gpu"): the oracle reproduces every tensor of the generator's and the critic's state_dict -> the
    oracle stays pinned where /root/reference is absent;
  * GPU (-m gpu): the native step plans reproduce them directly (native vs reference, no oracle in between).
Tolerances in units of the Adam step (lr = 1e-3): an element whose gradient is at rounding level may move by +-lr per
step in either direction (BN-shadowed biases have an analytically zero gradient and do nothing else), everything else
agrees to a few percent of a step on average."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import tabular_countergan as T

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LR = 1e-3


def _gold(name):
    z = np.load(os.path.join(GOLD, f"tabular_{name}.npz"))
    B, steps, seed0 = (int(v) for v in z["meta"])
    get = lambda p: OrderedDict((k[len(p):], torch.from_numpy(z[k])) for k in z.files if k.startswith(p))  # noqa: E731
    return B, steps, seed0, get("G."), get("D.")


def _bn_shadowed(k):
    return k in ("net.0.bias", "net.3.bias", "net.6.bias") or k.endswith("fc1.bias") or k.endswith("fc2.bias")


def _close(ref_sd, mine, steps, tag, mean_tol):
    for k, v in ref_sd.items():
        if "num_batches" in k:
            continue
        m, v = mine[k].detach().float().cpu(), v.float()
        if tag == "G" and _bn_shadowed(k):
            assert (v - m).abs().max() <= 2.02 * LR * steps, (tag, k)
        elif "running" in k or k.endswith("_u") or k.endswith("_v"):
            assert torch.allclose(v, m, atol=2.02 * LR * steps, rtol=5e-3), (tag, k, (v - m).abs().max())
        else:
            d = (v - m).abs()
            assert d.max() <= 2.02 * LR * steps and d.mean() <= mean_tol * LR, (tag, k, d.max().item(), d.mean().item())


def _moons_state():
    gs, ds, cs = T.moons_shapes()
    return gs, T.synth_params(gs, 1), T.synth_params(ds, 2), T.synth_params(cs, 3), T.sn_buffers(T.moons_d_dims(), 4)


def _kc_state():
    gs, ds, cs = T.kc_shapes()
    return (gs, T.synth_params(gs, 1), T.synth_params(ds, 2), T.synth_params(cs, 3), T.sn_buffers(T.kc_d_dims(), 4),
            T.bn_buffers(cs, 5, randomize=True))


def test_oracle_reproduces_moons_golden():
    B, steps, seed0, G1, D1 = _gold("moons")
    gs, PG, PD, PC, BD = _moons_state()
    S = T.make_state(PG, T.bn_buffers(gs), PD, BD, PC)
    for i in range(steps):
        T.moons_step(S, *T.moons_batch(B, seed0 + i))
    _close(G1, {**S["G"], **S["GB"]}, steps, "G", 0.02)
    _close(D1, {**S["D"], **S["DB"]}, steps, "D", 0.02)


def test_oracle_reproduces_kc_golden():
    B, steps, seed0, G1, D1 = _gold("kc")
    gs, PG, PD, PC, BD, BC = _kc_state()
    S = T.make_state(PG, T.bn_buffers(gs), PD, BD, PC, BC)
    nv = T.kc_norm_vals()
    for i in range(steps):
        T.kc_step(S, *T.kc_batch(B, seed0 + i), nv)
    _close(G1, {**S["G"], **S["GB"]}, steps, "G", 0.02)
    _close(D1, {**S["D"], **S["DB"]}, steps, "D", 0.02)


def _load_critic(plan_D, PD, BD):
    plan_D.flat.load(PD)
    for i, L in enumerate(plan_D.layers):
        L.u.copy_(BD[f"net.{2 * i}.weight_u"])
        L.v.copy_(BD[f"net.{2 * i}.weight_v"])


def _native_params(plan, G1, D1):
    g = {k: plan.G.p(k) for k in G1 if k in plan.G.shapes}
    d = {k: plan.D.flat.p(k) for k in D1 if k in plan.D.flat.shapes}
    return g, d


@pytest.mark.gpu
def test_native_moons_plan_reproduces_reference_golden():
    import pcg_b200  # noqa: F401
    from pcg_b200.tabular.moons import MoonsPlan
    B, steps, seed0, G1, D1 = _gold("moons")
    gs, PG, PD, PC, BD = _moons_state()
    plan = MoonsPlan(B, "cuda", use_graph=False)
    plan.G.load(PG)
    plan.C.load(PC)
    _load_critic(plan.D, PD, BD)
    plan.refresh()
    for i in range(steps):
        plan.step(*[t.cuda() for t in T.moons_batch(B, seed0 + i)])
    torch.cuda.synchronize()
    g, d = _native_params(plan, G1, D1)
    _close(OrderedDict((k, G1[k]) for k in g), g, steps, "G", 0.1)
    _close(OrderedDict((k, D1[k]) for k in d), d, steps, "D", 0.1)
    assert len(g) >= 14 and len(d) == 8


@pytest.mark.gpu
def test_native_kc_plan_reproduces_reference_golden():
    import pcg_b200  # noqa: F401
    from pcg_b200.tabular.kc import KcPlan
    B, steps, seed0, G1, D1 = _gold("kc")
    gs, PG, PD, PC, BD, BC = _kc_state()
    cat = OrderedDict((f, {"n": n, "raw_values": T.KC_RAW[f]}) for f, n in T.KC_CAT.items())
    plan = KcPlan(B, "cuda", cat, T.KC_CONT, use_graph=False)
    plan.G.load(PG)
    plan.C.load(PC)
    for j, nm in enumerate(plan.c_bn_names):
        plan.c_rm[j].copy_(BC[nm + ".running_mean"])
        plan.c_rv[j].copy_(BC[nm + ".running_var"])
    _load_critic(plan.D, PD, BD)
    plan.refresh()
    for i in range(steps):
        x, y, t, mask, noise = T.kc_batch(B, seed0 + i)
        plan.step(x.cuda(), y.cuda(), t.cuda(), mask.cuda(), [e.cuda() for e in noise])
    torch.cuda.synchronize()
    g, d = _native_params(plan, G1, D1)
    _close(OrderedDict((k, G1[k]) for k in g), g, steps, "G", 0.1)
    _close(OrderedDict((k, D1[k]) for k in d), d, steps, "D", 0.1)
    assert len(g) >= 40 and len(d) == 8
