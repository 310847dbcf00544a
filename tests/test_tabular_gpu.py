"""GPU parity of the native tabular CounteRGAN plans (moons, KC house sales) against the oracle."""
from collections import OrderedDict

import pytest
import torch

from oracle import tabular_countergan as T

pytestmark = pytest.mark.gpu


def l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def _load_critic(plan_D, PD, BD):
    plan_D.flat.load(PD)
    for i, L in enumerate(plan_D.layers):
        L.u.copy_(BD[f"net.{2 * i}.weight_u"])
        L.v.copy_(BD[f"net.{2 * i}.weight_v"])


@pytest.mark.parametrize("graph", [False, True])
def test_moons_step_matches_oracle(graph):
    import pcg_b200  # noqa: F401
    from pcg_b200.tabular.moons import MoonsPlan
    gs, ds, cs = T.moons_shapes()
    PG, PD, PC = T.synth_params(gs, 1), T.synth_params(ds, 2), T.synth_params(cs, 3)
    BD = T.sn_buffers(T.moons_d_dims(), 4)
    S = T.make_state(PG, T.bn_buffers(gs), PD, BD, PC)
    B = 64
    plan = MoonsPlan(B, "cuda", use_graph=graph)
    plan.G.load(PG)
    plan.C.load(PC)
    _load_critic(plan.D, PD, BD)
    plan.refresh()
    names = ["d_loss", "g_loss", "g_adv", "g_cls", "l1", "l2", "mask_pen"]
    for step in range(3):
        b = T.moons_batch(B, 50 + step)
        sc, gr = T.moons_step(S, *b)
        got = plan.step(*[t.cuda() for t in b]).tolist()
        tol = 2e-4 if step == 0 else 3e-2
        for i, k in enumerate(names):
            assert abs(got[i] - sc[k]) <= tol * max(abs(sc[k]), 0.1), (step, k, got[i], sc[k])
        if step == 0:
            for k in gr["G"]:
                if k in ("net.0.bias", "net.3.bias", "net.6.bias"):
                    continue        # BN-shadowed biases: analytically zero gradient
                assert l2(plan.G.g(k), gr["G"][k]) < 2e-3, (k, l2(plan.G.g(k), gr["G"][k]))
    # spectral-norm vectors advanced three times per step on both sides
    assert l2(plan.D.layers[0].u, S["DB"]["net.0.weight_u"]) < 1e-3
    assert l2(plan.gbn[1].rm, S["GB"]["net.4.running_mean"]) < 5e-2
    for k in S["D"]:
        assert ((plan.D.flat.p(k).cpu() - S["D"][k].detach()).abs().mean() / 1e-3).item() < 0.1, k


def test_moons_mirror_and_trainer(tmp_path):
    import numpy as np
    import pcg_b200  # noqa: F401
    from pcg_b200.tabular import moons as MO
    torch.manual_seed(0)
    G = MO.ResidualGenerator(2, 32, 3).cuda()
    C = MO.NNClassifier(2).cuda()
    assert list(G.state_dict().keys())[:7] == ["net.0.weight", "net.0.bias", "net.1.weight", "net.1.bias",
                                               "net.1.running_mean", "net.1.running_var", "net.1.num_batches_tracked"]
    x, y, t, mask = T.moons_batch(32, 7)
    PG = OrderedDict((k, v.detach().cpu()) for k, v in G.named_parameters())
    BG = OrderedDict((k, v.detach().cpu().clone()) for k, v in G.named_buffers())
    oh = torch.nn.functional.one_hot(t, 3).float()
    raw, masked = G(x.cuda(), oh.cuda(), mask.cuda())
    raw_o, masked_o = T.moons_g_forward(PG, BG, x, oh, mask)
    assert l2(raw, raw_o) < 1e-5 and l2(masked, masked_o) < 1e-5
    cfg = {"cuda": "cuda", "seed": 1, "epochs": 2, "batch_size": 32, "lr_G": 1e-3, "lr_D": 1e-3, "lambda_cls": 2.0,
           "lambda_reg_l1": 5.0, "lambda_reg_l2": 5.0, "lambda_mask": 3.0, "input_dim": 2, "hidden_dim": 32,
           "out_dir": str(tmp_path), "generator_path": str(tmp_path / "g.pt")}
    X = np.random.RandomState(0).randn(128, 2).astype(np.float32)
    yv = np.random.RandomState(1).randint(0, 3, 128)
    d, g = MO.train_countergan(G, cfg, X, yv, C)
    assert len(d) == 2 and np.isfinite(d).all() and np.isfinite(g).all()
    sd = torch.load(cfg["generator_path"])
    assert list(sd.keys()) == list(G.state_dict().keys())


def _kc_plan(B, graph):
    import pcg_b200  # noqa: F401
    from pcg_b200.tabular.kc import KcPlan
    cat = OrderedDict((f, {"n": n, "raw_values": T.KC_RAW[f]}) for f, n in T.KC_CAT.items())
    return KcPlan(B, "cuda", cat, T.KC_CONT, use_graph=graph)


@pytest.mark.parametrize("graph,B", [(False, 64), (True, 256), (True, 4096)])      # 4096 = BASELINE configs[2]
def test_kc_step_matches_oracle(graph, B):
    gs, ds, cs = T.kc_shapes()
    PG, PD, PC = T.synth_params(gs, 1), T.synth_params(ds, 2), T.synth_params(cs, 3)
    BD, BC = T.sn_buffers(T.kc_d_dims(), 4), T.bn_buffers(cs, 5, randomize=True)
    S = T.make_state(PG, T.bn_buffers(gs), PD, BD, PC, BC)
    plan = _kc_plan(B, graph)
    plan.G.load(PG)
    plan.C.load(PC)
    for j, nm in enumerate(plan.c_bn_names):
        plan.c_rm[j].copy_(BC[nm + ".running_mean"])
        plan.c_rv[j].copy_(BC[nm + ".running_var"])
    _load_critic(plan.D, PD, BD)
    plan.refresh()
    nv = T.kc_norm_vals()
    names = ["d_loss", "g_loss", "g_adv", "g_cls", "reg", "mask_pen"]
    for step in range(2):
        b = T.kc_batch(B, 80 + step)
        sc, gr = T.kc_step(S, *b, nv)
        x, y, t, mask, noise = b
        got = plan.step(x.cuda(), y.cuda(), t.cuda(), mask.cuda(), [e.cuda() for e in noise]).tolist()
        # the adversarial terms are differences of O(0.1) critic outputs and (from g_adv on) depend on the Adam update of
        # the critic, whose +-lr steps flip for gradients at rounding level: compare on the scale of the critic outputs
        tol = 2e-4 if step == 0 else 3e-2
        for i, k in enumerate(names):
            assert abs(got[i] - sc[k]) <= tol * max(abs(sc[k]), 0.1), (step, k, got[i], sc[k])
        if step == 0:
            assert l2(plan.xcf, gr["x_cf"]) < 1e-5
            for k in gr["G"]:
                if k.endswith("fc1.bias") or k.endswith("fc2.bias") or gr["G"][k] is None:
                    continue        # BN-shadowed biases: analytically zero gradient
                assert l2(plan.G.g(k), gr["G"][k]) < 3e-3, (k, l2(plan.G.g(k), gr["G"][k]))
            d = plan.diagnostics()
            for k in ("pred_gain", "sparsity", "l2", "flip"):
                assert abs(d[k] - sc[k]) <= 2e-3 * max(abs(sc[k]), 1e-2), (k, d[k], sc[k])
    for k in S["D"]:
        assert ((plan.D.flat.p(k).cpu() - S["D"][k].detach()).abs().mean() / 1e-3).item() < 0.1, k


def test_kc_mirror_and_trainer(tmp_path):
    import numpy as np
    import pcg_b200  # noqa: F401
    from pcg_b200.tabular import kc as KC
    cat = OrderedDict((f, {"n": n, "raw_values": T.KC_RAW[f]}) for f, n in T.KC_CAT.items())
    torch.manual_seed(0)
    G = KC.ResidualGenerator(17, 32, 4, T.KC_CONT, cat).cuda()
    C = KC.NNClassifier(17, 4).cuda().eval()
    gs, _, cs = T.kc_shapes()
    assert [k for k, _ in G.named_parameters()] == list(gs.keys())
    assert [k for k, _ in C.named_parameters()] == list(cs.keys())
    x, y, t, mask, noise = T.kc_batch(32, 3)
    cont, logits, samples = G(x.cuda(), torch.nn.functional.one_hot(t, 4).float().cuda(), mask.cuda())
    assert cont.shape == (32, 10) and samples[1].shape == (32, 30) and abs(samples[8].sum(1).mean().item() - 1) < 1e-5
    cfg = {"cuda": "cuda", "seed": 1, "epochs": 1, "batch_size": 64, "lr_G": 1e-3, "lr_D": 1e-3, "lambda_cls": 2.0,
           "lambda_reg": 1.0, "lambda_mask": 1.0, "input_dim": 17, "hidden_dim": 32, "gumbel_tau": 0.5, "scaler": None,
           "categorical_info": cat, "continuous_idx": T.KC_CONT, "immutable_idx": T.KC_IMMUTABLE,
           "out_dir": str(tmp_path), "generator_path": str(tmp_path / "g.pt")}
    X = torch.cat([T.kc_batch(64, 200 + i)[0] for i in range(3)]).numpy()
    yv = np.random.RandomState(1).randint(0, 4, len(X))
    d, g = KC.train_countergan(G, cfg, X, yv, C)
    assert np.isfinite(d).all() and np.isfinite(g).all()
    assert list(torch.load(cfg["generator_path"]).keys()) == list(G.state_dict().keys())


def test_capture_survives_cyclic_garbage():
    """Regression (round-1 GPUTEST failure): a dead plan is a reference cycle (plan <-> GraphStep) that holds a
    ``torch.cuda.CUDAGraph``; if Python's cyclic collector frees it while another plan is capturing, the graph's
    destructor (cudaGraphExecDestroy) invalidates that capture.  ``pcg_b200.graphs.capture`` collects before and keeps
    the collector off during the capture."""
    import gc
    import pcg_b200  # noqa: F401
    from pcg_b200.tabular.moons import MoonsPlan
    b = [t.cuda() for t in T.moons_batch(32, 1)]
    gc.collect()
    old = gc.get_threshold()
    gc.disable()
    try:
        dead = MoonsPlan(32, "cuda", use_graph=True)
        dead.step(*b)                      # captured: the plan now owns a CUDAGraph
        torch.cuda.synchronize()
        del dead                           # unreachable, but only the cyclic collector can free it
        gc.enable()
        gc.set_threshold(1, 1, 1)          # a collection at (nearly) every allocation, also inside the capture
        plan = MoonsPlan(32, "cuda", use_graph=True)
        gs, ds, cs = T.moons_shapes()
        plan.G.load(T.synth_params(gs, 1))
        plan.C.load(T.synth_params(cs, 3))
        _load_critic(plan.D, T.synth_params(ds, 2), T.sn_buffers(T.moons_d_dims(), 4))
        plan.refresh()
        sc = plan.step(*b).clone()
        sc2 = plan.step(*b).clone()
        torch.cuda.synchronize()
        assert torch.isfinite(sc).all() and torch.isfinite(sc2).all()
    finally:
        gc.set_threshold(*old)
        gc.enable()
