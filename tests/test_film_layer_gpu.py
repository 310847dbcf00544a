"""csrc/film_layer.cu: one call (two launches) per half of a FiLM residual block (house_sales_kc_usa/models/generator.py:19-35),
against float64 torch autograd and against the primitive operators it replaces; and the KC step plan with the fused
layers against the same plan on primitive operators."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


@pytest.mark.parametrize("M,H", [(4096, 32), (1000, 32), (64, 32), (7, 32), (2048, 64), (333, 64), (20011, 32)])
@pytest.mark.parametrize("relu", [True, False])
def test_fused_half_block_forward_and_backward(M, H, relu):
    import pcg_b200  # noqa: F401
    from pcg_b200 import ops as K
    assert K.film_layer_supported(M, H) and not K.film_layer_supported(64, 48)
    torch.manual_seed(M + H)
    dev = "cuda"
    x = torch.randn(M, H, device=dev)
    W = torch.randn(H, H, device=dev) * H ** -0.5
    bias, gam, bet = torch.randn(H, device=dev) * 0.1, 1 + 0.1 * torch.randn(H, device=dev), 0.1 * torch.randn(H, device=dev)
    fg, fb, res = 1 + 0.3 * torch.randn(M, H, device=dev), 0.3 * torch.randn(M, H, device=dev), torch.randn(M, H, device=dev)
    rm, rv = torch.randn(H, device=dev) * 0.1, 1 + 0.1 * torch.rand(H, device=dev)
    rm0, rv0 = rm.clone(), rv.clone()
    nbt = torch.zeros((), dtype=torch.int64, device=dev)
    st = K.BNState(H, dev)
    u, n, out = (torch.full((M, H), 9.0, device=dev) for _ in range(3))
    K.film_layer_fwd(x, W, bias, gam, bet, rm, rv, nbt, st, fg, fb, u, n, out, res=None if relu else res, relu=relu)
    xd, Wd, bd, gd, btd = (t.double().requires_grad_(True) for t in (x, W, bias, gam, bet))
    fgd, fbd = fg.double().requires_grad_(True), fb.double().requires_grad_(True)
    ud = xd @ Wd.t() + bd
    mu, var = ud.mean(0), ud.var(0, unbiased=False)
    nd = (ud - mu) * torch.rsqrt(var + 1e-5) * gd + btd
    fd = fgd * nd + fbd
    outd = torch.relu(fd) if relu else res.double() + fd
    assert rel(u, ud) < 1e-5 and rel(n, nd) < 2e-5 and rel(out, outd) < 2e-5
    assert rel(st.mean, mu) < 1e-5 and rel(st.rstd, torch.rsqrt(var + 1e-5)) < 1e-5 and nbt.item() == 1
    if M > 1:
        assert rel(rm, 0.9 * rm0.double() + 0.1 * mu) < 1e-5
        assert rel(rv, 0.9 * rv0.double() + 0.1 * ud.var(0, unbiased=True)) < 1e-5
    # backward: cotangent d_f of f (the caller applies relu' / passes the residual gradient through)
    d_f = torch.randn(M, H, device=dev)
    skip, ref = torch.randn(M, H, device=dev), torch.randn(M, H, device=dev)
    gx, gW, gfg, gfb, gg, gb = torch.autograd.grad(fd, [xd, Wd, fgd, fbd, gd, btd], d_f.double())
    dfg, dfb = torch.ones(M, H, device=dev), torch.ones(M, H, device=dev)
    du, dx, dgam, dbet = torch.empty(M, H, device=dev), torch.empty(M, H, device=dev), torch.empty(H, device=dev), torch.empty(H, device=dev)
    K.film_layer_bwd(d_f, fg, n, u, st, gam, W, dfg, dfb, du, dx, dgam, dbet, accumulate=True)
    tol = 2e-4 if M > 16 else 2e-3
    assert rel(dfg - 1, gfg) < tol and rel(dfb - 1, gfb) < tol
    assert rel(dx, gx) < tol and rel(dgam, gg) < tol and rel(dbet, gb) < tol
    assert rel(du.double().t() @ x.double(), gW) < tol                      # du is what the weight gradient consumes
    K.film_layer_bwd(d_f, fg, n, u, st, gam, W, dfg, dfb, du, dx, dgam, dbet, add_src=skip, act_ref=ref)
    assert rel(dfg, gfg) < tol and rel(dx, (gx + skip.double()) * (ref > 0)) < tol


@pytest.mark.parametrize("B", [256, 4096])
def test_kc_plan_fused_layers_match_primitive_operators(B, monkeypatch):
    """The KC CounteRGAN iteration with the fused half blocks against the same plan composed from primitive operators:
    every generator gradient and the updated parameters after two iterations."""
    import pcg_b200  # noqa: F401
    from collections import OrderedDict
    from oracle import tabular_countergan as T
    from pcg_b200.tabular.kc import KcPlan
    gs, ds, cs = T.kc_shapes()
    PG, PD, PC = T.synth_params(gs, 1), T.synth_params(ds, 2), T.synth_params(cs, 3)
    BD, BC = T.sn_buffers(T.kc_d_dims(), 4), T.bn_buffers(cs, 5, randomize=True)
    cat = OrderedDict((f, {"n": n, "raw_values": T.KC_RAW[f]}) for f, n in T.KC_CAT.items())
    # the Linear biases in front of a BatchNorm have an analytically zero gradient: rounding noise on both sides
    shadowed = lambda k: k.endswith("fc1.bias") or k.endswith("fc2.bias")  # noqa: E731
    plans = []
    for fused in ("1", "0"):
        monkeypatch.setenv("PCG_FILM_LAYER", fused)
        plan = KcPlan(B, "cuda", cat, T.KC_CONT)
        assert plan.fused == (fused == "1")
        plan.G.load(PG); plan.C.load(PC); plan.D.flat.load(PD)
        for j, nm in enumerate(plan.c_bn_names):
            plan.c_rm[j].copy_(BC[nm + ".running_mean"]); plan.c_rv[j].copy_(BC[nm + ".running_var"])
        for i, L in enumerate(plan.D.layers):
            L.u.copy_(BD[f"net.{2 * i}.weight_u"]); L.v.copy_(BD[f"net.{2 * i}.weight_v"])
        plan.refresh()
        plans.append(plan)
    for it in range(2):
        x, y, t, m, noise = T.kc_batch(B, 80 + it)
        outs = [p.step(x.cuda(), y.cuda(), t.cuda(), m.cuda(), [e.cuda() for e in noise]).clone() for p in plans]
        assert rel(outs[0][:6], outs[1][:6]) < 1e-4, (it, outs)
        for k in plans[0].G.names:
            g0, g1 = plans[0].G.g(k), plans[1].G.g(k)
            if not shadowed(k) and g1.abs().max() > 0:
                assert rel(g0, g1) < 2e-3, (it, k)
    for k in plans[0].G.names:
        d = (plans[0].G.p(k) - plans[1].G.p(k)).abs()
        assert d.max() <= 2.02e-3 * 2, (k, d.max())
        if not shadowed(k):
            assert d.mean() <= 0.05e-3, (k, d.max(), d.mean())
    for b0, b1 in zip(plans[0].blk, plans[1].blk):
        for nm in ("bn1", "bn2"):
            # the running mean follows the (free-walking, see above) Linear bias: +-lr per step on either side
            assert (b0[nm].rm - b1[nm].rm).abs().max() < 5e-4 and rel(b0[nm].rv, b1[nm].rv) < 1e-4


@pytest.mark.parametrize("M,H,n", [(4096, 32, 10), (777, 32, 3), (64, 64, 4), (300, 32, 1)])
def test_chained_half_blocks_equal_single_calls(M, H, n):
    """pcg_film_chain_fwd (n + 1 launches) against n calls of pcg_film_layer_fwd: same outputs, saved statistics and running
    buffers (the two compute the same sums in the same order: bit-identical)."""
    import pcg_b200  # noqa: F401
    from pcg_b200 import ops as K
    torch.manual_seed(M + H + n)
    dev = "cuda"
    x = torch.randn(M, H, device=dev)

    def make():
        torch.manual_seed(5)
        hs = []
        for k in range(n):
            W = torch.randn(H, H, device=dev) * H ** -0.5
            bias, gam, bet = torch.randn(H, device=dev) * 0.1, 1 + 0.1 * torch.randn(H, device=dev), 0.1 * torch.randn(H, device=dev)
            fg, fb = 1 + 0.3 * torch.randn(M, H, device=dev), 0.3 * torch.randn(M, H, device=dev)
            rm, rv = torch.zeros(H, device=dev), torch.ones(H, device=dev)
            nbt = torch.zeros((), dtype=torch.int64, device=dev)
            u, nn_, out = (torch.full((M, H), 9.0, device=dev) for _ in range(3))
            hs.append([W, bias, gam, bet, rm, rv, nbt, K.BNState(H, dev), fg, fb, None, u, nn_, out])
        return hs
    a, b = make(), make()
    # odd half blocks are residual: res = the input of the previous (relu) half block
    cur = x
    for k, h in enumerate(a):
        res = None if k % 2 == 0 else (x if k == 1 else a[k - 2][13])
        h[10] = res
        K.film_layer_fwd(cur, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[11], h[12], h[13], res=res,
                         relu=res is None)
        cur = h[13]
    for k, h in enumerate(b):
        h[10] = None if k % 2 == 0 else (x if k == 1 else b[k - 2][13])
    K.film_chain_fwd(x, [tuple(h) for h in b])
    for k, (ha, hb) in enumerate(zip(a, b)):
        for idx in (11, 12, 13, 4, 5):
            assert torch.equal(ha[idx], hb[idx]), (k, idx)
        assert torch.equal(ha[7].mean, hb[7].mean) and torch.equal(ha[7].rstd, hb[7].rstd) and ha[6].item() == hb[6].item() == 1
