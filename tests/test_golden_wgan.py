"""Golden vectors of the conditional WGAN-GP iteration produced by the reference script's own classes and loop body
(tests/golden/make_golden_wgan.py; widths 64/64/32, batch 8, n_critic 2, 3 iterations):
  * CPU (-m "not gpu"): the oracle reproduces them -> oracle/wgan_gp.py stays pinned where /root/reference is absent;
  * GPU (-m gpu): the native plan (exact fp32 CUDA-core mode) reproduces them directly.
Per tensor the fixture holds [sum, abs-sum, 32 samples]; samples are compared in units of the AdamW step (lr = 1e-4,
beta1 = 0: an element moves by up to sqrt(10) * lr per step along its gradient's sign, and an element with a rounding-level
gradient does so on either side)."""
import os

import numpy as np
import pytest
import torch

from oracle import wgan_gp as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wgan_gp.npz")
LR = 1e-4
CFG_KEYS = ("batchsize", "critic_hidden_size", "critic_size", "generator_size", "latent_size", "n_critic")


def _summary(t):
    t = t.detach().double().flatten().cpu()
    idx = torch.linspace(0, t.numel() - 1, 32).long()
    return np.concatenate([[t.sum().item(), t.abs().sum().item()], t[idx].numpy()])


def _meta(z):
    m = [int(v) for v in z["meta"]]
    steps, seed0, seed_g, seed_c = m[:4]
    return steps, seed0, seed_g, seed_c, O.Hyper(**dict(zip(CFG_KEYS, m[4:])))


def _check(z, get, steps, med_tol, abs_tol):
    for key in z.files:
        if key in ("meta", "losses"):
            continue
        net, k = key[0], key[2:]
        if "num_batches" in k:
            continue
        ref, got = z[key], _summary(get(net, k))
        if "running" in k:
            assert np.allclose(got[2:], ref[2:], rtol=5e-3, atol=5e-4), key
            continue
        d = np.abs(got[2:] - ref[2:])
        assert d.max() <= 2 * 3.2 * LR * steps + 1e-7, (key, d.max())
        if k in O.SHADOWED:
            continue                     # zero-gradient biases in front of a norm layer: free random walk on both sides
        assert np.median(d) <= med_tol * LR + 1e-7, (key, np.median(d))
        assert abs(got[1] - ref[1]) <= abs_tol * abs(ref[1]) + 1e-5, key


def _check_losses(z, it, sc, tol):
    ref = z["losses"][it]
    assert abs(sc["critic_loss"] - ref[0]) < tol * abs(ref[0]) + 1e-5, (it, sc, ref)
    assert abs(sc["gp"] - ref[1]) < tol * abs(ref[1]) + 1e-5, (it, sc, ref)
    if not np.isnan(ref[2]):
        assert abs(sc["generator_loss"] - ref[2]) < tol * abs(ref[2]) + 1e-5, (it, sc, ref)


def test_oracle_reproduces_reference_golden():
    z = np.load(GOLD)
    steps, seed0, seed_g, seed_c, hp = _meta(z)
    S = O.make_state(O.synth_params(O.g_shapes(hp), seed_g), O.g_buffers(hp), O.synth_params(O.c_shapes(hp), seed_c))
    eye = torch.eye(hp.num_classes)
    for it in range(steps):
        b = O.synth_batch(hp, hp.batchsize, seed0 + it)
        sc, _ = O.critic_step(S, hp, b["real"], eye[b["labels"]], b["noise"], b["alpha"])
        if it % hp.n_critic == 0:
            sc.update(O.generator_step(S, hp, b["noise_g"], eye[b["labels_g"]])[0])
        _check_losses(z, it, sc, 2e-5 if it == 0 else 2e-3)
    get = lambda net, k: (S[net][k] if k in S[net] else S["GB"][k])  # noqa: E731
    _check(z, get, steps, 0.05, 3e-4)


def test_module_mirrors_have_the_reference_state_dict():
    """pcg_b200.wgan.Generator / Critic (constructed on CPU, no forward) carry exactly the tensors of the reference classes
    (names and shapes as restated by the oracle, which tests/test_wgan_oracle_vs_reference.py pins to the lifted classes)."""
    import pcg_b200  # noqa: F401
    from pcg_b200.wgan import Critic, Generator, Hyperparameter
    from pcg_b200.wgan import plan as W
    for cfg in (dict(), dict(critic_size=64, generator_size=128, critic_hidden_size=32, latent_size=8)):
        hp, ohp = Hyperparameter(**cfg), O.Hyper(**cfg)
        g = {k: tuple(v.shape) for k, v in Generator(hp).state_dict().items() if "running" not in k and "num_batches" not in k}
        c = {k: tuple(v.shape) for k, v in Critic(hp).state_dict().items()}
        assert list(g.items()) == list(O.g_shapes(ohp).items()) == list(W.g_shapes(hp).items())
        assert list(c.items()) == list(O.c_shapes(ohp).items()) == list(W.c_shapes(hp).items())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Critic(Hyperparameter(critic_size=64, critic_hidden_size=32))(torch.zeros(2, 1, 28, 28), torch.zeros(2, 10))


@pytest.mark.gpu
def test_native_plan_reproduces_reference_golden():
    import pcg_b200  # noqa: F401
    from pcg_b200.wgan import WganGpPlan
    z = np.load(GOLD)
    steps, seed0, seed_g, seed_c, hp = _meta(z)
    plan = WganGpPlan(hp, hp.batchsize, "cuda", use_graph=True, tensor_cores=False)
    plan.G.load(O.synth_params(O.g_shapes(hp), seed_g))
    plan.C.load(O.synth_params(O.c_shapes(hp), seed_c))
    plan.refresh()
    for it in range(steps):
        b = {k: v.cuda() for k, v in O.synth_batch(hp, hp.batchsize, seed0 + it).items()}
        if it % hp.n_critic == 0:
            s = plan.step(b["real"], b["labels"], b["noise"], b["alpha"], b["labels_g"], b["noise_g"]).tolist()
        else:
            s = plan.step(b["real"], b["labels"], b["noise"], b["alpha"]).tolist()
        _check_losses(z, it, {"critic_loss": s[0], "gp": s[3], "generator_loss": s[4]}, 1e-4 if it == 0 else 5e-3)

    def get(net, k):
        arena = plan.G if net == "G" else plan.C
        if k in arena.shapes:
            return arena.p(k)
        i = ["tcnn.1", "tcnn.4", "tcnn.7"].index(k.rsplit(".", 1)[0])
        return plan.g_bn[i][{"running_mean": "rm", "running_var": "rv"}[k.rsplit(".", 1)[1]]]
    _check(z, get, steps, 0.1, 1e-3)
