"""GPU parity of the native moons MLP GAN plans (conditional and unconditional) against the oracle."""
import pytest
import torch

from oracle import moons_gan as M

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


@pytest.mark.parametrize("label_dim,B,graph,fused", [(2, 1024, True, False), (0, 128, True, False), (2, 50, False, False),
                                                     (2, 1024, True, True), (0, 128, True, True), (2, 50, True, True),
                                                     (2, 777, True, True), (0, 7, True, True)])
def test_gan_step_matches_oracle(label_dim, B, graph, fused):
    """fused = the whole iteration in one cluster launch (pcg_mlp_gan_step); else the operator-composed plan."""
    import pcg_b200  # noqa: F401
    from pcg_b200.moons.gan import MlpGanPlan
    gs, ds = M.shapes(label_dim=label_dim)
    PG, PD = M.synth_params(gs, 1), M.synth_params(ds, 2)
    S = M.make_state(PG, PD)
    plan = MlpGanPlan(B, 32, label_dim, 128, "cuda", use_graph=graph, fused=fused)
    plan.G.load({"net." + k: v for k, v in PG.items()})
    plan.D.load({"net." + k: v for k, v in PD.items()})
    plan.refresh()
    cu = lambda t: None if t is None else t.cuda().contiguous()  # noqa: E731
    for step in range(3):
        b = M.synth_batch(B, 300 + step, label_dim=label_dim)
        pG0 = {k: v.detach().clone() for k, v in S["G"].items()}
        ld, lg, gr = M.gan_step(S, *b)
        sc = plan.step(*[cu(t) for t in b]).tolist()
        tol = 2e-5 if step == 0 else 2e-3
        assert abs(sc[0] - ld) <= tol * 10 * abs(ld) and abs(sc[1] - lg) <= tol * 10 * abs(lg), (step, sc[:2], ld, lg)
        if step == 0:
            for k in gr["D"]:
                assert rel(plan.D.g("net." + k), gr["D"][k]) < 2e-4, k
            for k in gr["G"]:
                assert rel(plan.G.g("net." + k), gr["G"][k]) < 2e-4, k
            for k in pG0:       # Adam update in units of lr (robust mean, see test_mnist_step_gpu)
                d_nat = plan.G.p("net." + k).cpu() - pG0[k]
                d_or = S["G"][k].detach() - pG0[k]
                assert ((d_nat - d_or).abs().mean() / 1e-3).item() < 0.02, k


def test_mirror_modules_and_train_functions():
    import numpy as np
    import pcg_b200  # noqa: F401
    from pcg_b200.moons import gan as GAN
    torch.manual_seed(0)
    G, D = GAN.Generator(32, 2, 128).cuda(), GAN.Discriminator(2, 128).cuda()
    assert list(G.state_dict().keys()) == ["net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias"]
    z, oh = torch.randn(7, 32, device="cuda"), torch.eye(2, device="cuda")[torch.randint(0, 2, (7,), device="cuda")]
    P = {k.replace("net.", ""): v.detach().cpu() for k, v in G.state_dict().items()}
    assert rel(G(z, oh), M.g_forward(P, z.cpu(), oh.cpu())) < 1e-5
    PD = {k.replace("net.", ""): v.detach().cpu() for k, v in D.state_dict().items()}
    x = torch.randn(7, 2, device="cuda")
    assert rel(D(x, oh), M.d_forward(PD, x.cpu(), oh.cpu())) < 1e-5
    cfg = {"n_samples": 256, "z_dim": 32, "hidden_dim": 128, "label_dim": 2, "batch_size": 64, "lr": 1e-3, "epochs": 3}
    X = np.random.RandomState(0).randn(256, 2).astype(np.float32)
    lD, lG = GAN.train_cgan(torch.from_numpy(X), torch.randint(0, 2, (256,)), G, D, cfg)
    assert len(lD) == 3 and all(np.isfinite(lD)) and all(np.isfinite(lG))
    G2, D2 = GAN.build_generator(32, 128), GAN.build_discriminator(128)
    lD, lG = GAN.train_gan(X, G2, D2, cfg)
    assert len(lD) == 3 and all(np.isfinite(lD))
    with pytest.raises(ValueError):
        GAN.train_gan(X[:100], G2, D2, cfg)
