"""GPU: the data-flow captured graph (pcg_b200/dataflow.py: operators on up to 16 streams, ordered by their memory
hazards only) reproduces the sequentially captured graph BIT FOR BIT on the tabular CounteRGAN plans - scalars,
parameters, Adam state, BatchNorm buffers, spectral-norm vectors - over several iterations."""
from collections import OrderedDict

import pytest
import torch

from oracle import tabular_countergan as T

pytestmark = pytest.mark.gpu


def _critic(plan_D, PD, BD):
    plan_D.flat.load(PD)
    for i, L in enumerate(plan_D.layers):
        L.u.copy_(BD[f"net.{2 * i}.weight_u"])
        L.v.copy_(BD[f"net.{2 * i}.weight_v"])


def _run_kc(parallel, monkeypatch, B=512, steps=3):
    import pcg_b200  # noqa: F401
    from pcg_b200.tabular.kc import KcPlan
    monkeypatch.setenv("PCG_DATAFLOW", "1" if parallel else "0")
    gs, ds, cs = T.kc_shapes()
    PG, PD, PC = T.synth_params(gs, 1), T.synth_params(ds, 2), T.synth_params(cs, 3)
    BD, BC = T.sn_buffers(T.kc_d_dims(), 4), T.bn_buffers(cs, 5, randomize=True)
    cat = OrderedDict((f, {"n": n, "raw_values": T.KC_RAW[f]}) for f, n in T.KC_CAT.items())
    plan = KcPlan(B, "cuda", cat, T.KC_CONT, use_graph=True)
    plan.G.load(PG)
    plan.C.load(PC)
    for j, nm in enumerate(plan.c_bn_names):
        plan.c_rm[j].copy_(BC[nm + ".running_mean"])
        plan.c_rv[j].copy_(BC[nm + ".running_var"])
    _critic(plan.D, PD, BD)
    plan.refresh()
    out = []
    for s in range(steps):
        x, y, t, mask, noise = T.kc_batch(B, 80 + s)
        out.append(plan.step(x.cuda(), y.cuda(), t.cuda(), mask.cuda(), [e.cuda() for e in noise]).clone())
    torch.cuda.synchronize()
    state = [torch.stack(out), plan.G.data, plan.G.m, plan.G.v, plan.D.flat.data, plan.D.flat.m, plan.xcf]
    state += [b["bn1"].rm for b in plan.blk] + [b["bn2"].rv for b in plan.blk] + [L.u for L in plan.D.layers]
    return [t.clone() for t in state], plan


def _run_moons(parallel, monkeypatch, B=64, steps=4):
    import pcg_b200  # noqa: F401
    from pcg_b200.tabular.moons import MoonsPlan
    monkeypatch.setenv("PCG_DATAFLOW", "1" if parallel else "0")
    gs, ds, cs = T.moons_shapes()
    plan = MoonsPlan(B, "cuda", use_graph=True)
    plan.G.load(T.synth_params(gs, 1))
    plan.C.load(T.synth_params(cs, 3))
    _critic(plan.D, T.synth_params(ds, 2), T.sn_buffers(T.moons_d_dims(), 4))
    plan.refresh()
    out = []
    for s in range(steps):
        out.append(plan.step(*[t.cuda() for t in T.moons_batch(B, 50 + s)]).clone())
    torch.cuda.synchronize()
    state = [torch.stack(out), plan.G.data, plan.G.m, plan.D.flat.data, plan.D.flat.v, plan.xcf]
    state += [b.rm for b in plan.gbn] + [L.v for L in plan.D.layers]
    return [t.clone() for t in state], plan


@pytest.mark.parametrize("run", [_run_kc, _run_moons])
def test_dataflow_graph_is_bit_identical_to_the_sequential_graph(run, monkeypatch):
    seq, plan_s = run(False, monkeypatch)
    par, plan_p = run(True, monkeypatch)
    assert plan_s.run.program is None and plan_p.run.program is not None
    prog = plan_p.run.program
    n, depth = len(prog.ops), prog.critical_path()
    print(f"{run.__name__}: {n} operator launches, critical path {depth}, {prog.n_streams} streams")
    assert prog.n_streams > 4 and depth < 0.5 * n
    for a, b in zip(seq, par):
        assert torch.equal(a, b)
