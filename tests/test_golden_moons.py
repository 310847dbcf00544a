"""Golden vectors of the two moons MLP GANs produced by the reference itself (tests/golden/make_golden_moons.py: the
reference's own classes and training loop, AST-lifted and run unmodified with injected random draws):
  * CPU (-m "not gpu"): the oracle reproduces them -> oracle/moons_gan.py stays pinned where /root/reference is absent;
  * GPU (-m gpu): the one-launch kernel ``pcg_mlp_gan_step`` and the operator-composed plan reproduce them directly
    (native vs reference, no oracle in between).
Tolerances: parameters after n Adam steps in units of the step size (an element with a noise-level gradient may move
by up to lr per step in either direction): max <= 2.02 * lr * n, median <= 0.02 * lr; loss sums 1e-4 (oracle) / 2e-3."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import moons_gan as M

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LR = 1e-3


def _load(name):
    z = np.load(os.path.join(GOLD, f"moons_{name}.npz"))
    B, nb, seed0, label_dim, epochs = (int(v) for v in z["meta"])
    batches = [M.synth_batch(B, seed0 + i, label_dim=label_dim) for i in range(epochs * nb)]
    for e in range(1, epochs):                     # the simple GAN re-reads the same X every epoch (shuffle disabled)
        for i in range(nb):
            b = list(batches[e * nb + i])
            b[0] = batches[i][0]
            batches[e * nb + i] = tuple(b)
    get = lambda p: OrderedDict((k[len(p):], torch.from_numpy(z[k])) for k in z.files if k.startswith(p))  # noqa: E731
    return z, B, nb, label_dim, epochs, batches, get


def _check_params(got, ref, n_steps, what):
    for k, r in ref.items():
        d = (got[k].detach().float().cpu() - r).abs().numpy().ravel()
        assert d.max() <= 2.02 * LR * n_steps + 1e-7, (what, k, d.max())
        assert np.median(d) <= 0.02 * LR + 1e-7, (what, k, np.median(d))


@pytest.mark.parametrize("name", ["cgan", "gan"])
def test_oracle_reproduces_reference_goldens(name):
    z, B, nb, label_dim, epochs, batches, get = _load(name)
    S = M.make_state(get("G0."), get("D0."))
    lossD, lossG = [], []
    for e in range(epochs):
        d = g = 0.0
        for i in range(nb):
            ld, lg, _ = M.gan_step(S, *batches[e * nb + i])
            d, g = d + ld, g + lg
        lossD.append(d)
        lossG.append(g)
    assert np.allclose(lossD, z["loss_D"], rtol=1e-4) and np.allclose(lossG, z["loss_G"], rtol=1e-4)
    for net, pre in (("G", "G1."), ("D", "D1.")):
        for k, r in get(pre).items():
            assert torch.allclose(S[net][k].detach(), r, atol=2e-6, rtol=1e-4), (net, k)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", ["cgan", "gan"])
def test_native_reproduces_reference_goldens(name, fused):
    import pcg_b200  # noqa: F401
    from pcg_b200.moons.gan import MlpGanPlan
    z, B, nb, label_dim, epochs, batches, get = _load(name)
    plan = MlpGanPlan(B, 32, label_dim, 128, "cuda", lr=LR, fused=fused)
    plan.G.load({"net." + k: v for k, v in get("G0.").items()})
    plan.D.load({"net." + k: v for k, v in get("D0.").items()})
    plan.refresh()
    lossD, lossG = [], []
    for e in range(epochs):
        acc = torch.zeros(8, device="cuda")
        for i in range(nb):
            acc += plan.step(*[None if t is None else t.cuda().contiguous() for t in batches[e * nb + i]])
        tot = acc.tolist()
        lossD.append(tot[0])
        lossG.append(tot[1])
    assert np.allclose(lossD, z["loss_D"], rtol=2e-3) and np.allclose(lossG, z["loss_G"], rtol=2e-3)
    n_steps = epochs * nb
    _check_params({k: plan.G.p("net." + k) for k in get("G1.")}, get("G1."), n_steps, "G")
    _check_params({k: plan.D.p("net." + k) for k in get("D1.")}, get("D1."), n_steps, "D")
