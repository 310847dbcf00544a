"""Golden vectors produced by the reference itself (tests/golden/make_golden.py):
  * CPU (-m "not gpu"): the oracle reproduces them  -> the oracle stays pinned where /root/reference is absent;
  * GPU (-m gpu):       the native fp32 plan reproduces them directly -> native vs reference, no oracle in between.
"""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import mnist_countergan as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["small", "full"]


def _load(name):
    z = np.load(os.path.join(GOLD, f"mnist_{name}.npz"))
    ch, nres, B, n_steps = (int(v) for v in z["meta"])
    return z, ch, nres, B, n_steps


def _summary(t):
    t = t.detach().double().flatten().cpu()
    idx = torch.linspace(0, t.numel() - 1, 32).long()
    return np.concatenate([[t.sum().item(), t.abs().sum().item()], t[idx].numpy()])


def _params(ch, nres):
    return (O.synth_params(O.g_param_shapes(ch, nres), 1, "G"), O.g_buffers(ch, nres),
            O.synth_params(O.d_param_shapes(), 2, "D"), O.synth_params(O.c_param_shapes(), 3, "C"))


def _check_train_summaries(z, get, n_steps, lr_g, atol_lr):
    """get(net, key) -> tensor.  Summaries: [sum, abs-sum, 32 samples].  Samples are compared in units
    of the Adam step (an element with a noise-level gradient may differ by up to 2*lr per step)."""
    for key in z.files:
        if not key.startswith("train_"):
            continue
        net, k = key[6], key[8:]
        if "num_batches" in k:
            assert int(z[key][2]) == n_steps
            continue
        got, ref = _summary(get(net, k)), z[key]
        lr = lr_g if net == "G" else 1e-5
        if O.is_bn_shadowed_bias(k):
            assert np.abs(got[2:] - ref[2:]).max() <= 2.02 * lr * n_steps, key
            continue
        if "running" in k:
            # BN running statistics after n_steps (momentum 0.1): absolute 2e-5 on O(0.01-1) values
            assert np.allclose(got[2:], ref[2:], rtol=1e-3, atol=2e-5), key
            assert abs(got[1] - ref[1]) <= 1e-3 * abs(ref[1]) + 1e-4, key
            continue
        # samples: most agree to rounding; a few noise-gradient elements may differ by O(lr)
        d = np.abs(got[2:] - ref[2:])
        assert d.max() <= 2.02 * lr * n_steps + 1e-7, (key, d.max())
        assert np.median(d) <= atol_lr * lr + 1e-7, (key, np.median(d))
        assert abs(got[1] - ref[1]) <= 3e-4 * abs(ref[1]) + 1e-6, key


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_goldens(name):
    z, ch, nres, B, n_steps = _load(name)
    PG, BG, PD, PC = _params(ch, nres)
    x, y, t, mask = O.synth_batch(B, 700)
    with torch.no_grad():
        raw, masked = O.g_forward(PG, BG, x, t, mask, n_resblocks=nres)
        assert np.allclose(raw.numpy(), z["fwd_raw"], atol=1e-6, rtol=1e-4)
        assert np.allclose(masked.numpy(), z["fwd_masked"], atol=1e-6, rtol=1e-4)
        assert np.allclose(O.d_forward(PD, x, y).numpy(), z["fwd_d_logits"], atol=2e-6, rtol=1e-4)
        assert np.allclose(O.c_forward(PC, x).numpy(), z["fwd_c_logits"], atol=1e-5, rtol=1e-4)
        for k in z.files:
            if k.startswith("fwd_buf/"):
                assert np.allclose(BG[k[8:]].numpy(), z[k], atol=1e-6, rtol=1e-4), k
        raw_e, _ = O.g_forward(PG, BG, x, t, mask, n_resblocks=nres, training=False)
        assert np.allclose(raw_e.numpy(), z["fwd_raw_eval"], atol=1e-6, rtol=1e-4)
    PG, BG, PD, PC = _params(ch, nres)
    S = O.make_state(PG, BG, PD, PC)
    for i in range(n_steps):
        O.countergan_step(S, *O.synth_batch(B, 800 + i, mnist_like=(i % 2 == 1)), n_resblocks=nres)
    _check_train_summaries(z, lambda net, k: (S[net][k] if k in S[net] else S["GB"][k]), n_steps, 5e-5, 0.05)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_native_fp32_reproduces_reference_goldens(name):
    import pcg_b200  # noqa: F401
    from pcg_b200.mnist import plan as P
    z, ch, nres, B, n_steps = _load(name)
    PG, BG, PD, PC = _params(ch, nres)
    ga = P.Arena(0, ch, nres, "cuda").load_dict([v.cuda() for v in PG.values()])
    da = P.Arena(1, ch, nres, "cuda").load_dict([v.cuda() for v in PD.values()])
    ca = P.Arena(2, ch, nres, "cuda").load_dict([v.cuda() for v in PC.values()])
    running = torch.zeros(2 * nres, 2, ch, device="cuda")
    running[:, 1] = 1
    nbt = torch.zeros(2 * nres, dtype=torch.int64, device="cuda")
    plan = P.MnistStepPlan(B, ga, da, ca, running, nbt, P.StepConfig(precision="fp32"), ch, nres)
    x, y, t, mask = (v.cuda().contiguous() for v in O.synth_batch(B, 700))
    raw, masked = plan.g_forward(x, t, mask, True)
    assert np.allclose(raw.cpu().numpy(), z["fwd_raw"], atol=2e-6, rtol=2e-4)
    assert np.allclose(masked.cpu().numpy(), z["fwd_masked"], atol=2e-6, rtol=2e-4)
    assert np.allclose(plan.d_forward(x, y).cpu().numpy(), z["fwd_d_logits"], atol=1e-5, rtol=2e-4)
    assert np.allclose(plan.c_forward(x).cpu().numpy(), z["fwd_c_logits"], atol=2e-5, rtol=2e-4)
    raw_e, _ = plan.g_forward(x, t, mask, False)
    assert np.allclose(raw_e.cpu().numpy(), z["fwd_raw_eval"], atol=2e-6, rtol=2e-4)
    # training from fresh state
    ga.load_dict([v.cuda() for v in PG.values()])
    running.zero_()
    running[:, 1] = 1
    nbt.zero_()
    plan.refresh_weights()
    for i in range(n_steps):
        b = O.synth_batch(B, 800 + i, mnist_like=(i % 2 == 1))
        plan.step(*(v.cuda().contiguous() for v in b))
    torch.cuda.synchronize()
    gk, dk = list(O.g_param_shapes(ch, nres).items()), list(O.d_param_shapes().items())

    def get(net, k):
        if net == "G":
            names = [n for n, _ in gk]
            if k in names:
                i = names.index(k)
                return ga.view(i, gk[i][1])
            blk, bn, what = int(k.split(".")[1]), int(k.split(".")[2][2]), k.split(".")[3]
            row = 2 * blk + (bn - 1)
            if what == "num_batches_tracked":
                return nbt[row].float()
            return running[row, 0 if what == "running_mean" else 1]
        names = [n for n, _ in dk]
        i = names.index(k)
        return da.view(i, dk[i][1])
    _check_train_summaries(z, get, n_steps, 5e-5, 0.05)


@pytest.mark.gpu
def test_native_bf16_plan_against_reference_golden():
    """The BENCHMARKED precision (bf16 storage, tcgen05 convolutions) against the vectors the reference itself produced
    (full architecture): forwards to bf16 tolerance (relative L2 1e-2), and after the training iterations every sampled
    parameter within the Adam bound 2*lr*steps of the reference's, half of them within 0.5*lr, absolute sums to 1e-3."""
    import pcg_b200  # noqa: F401
    from pcg_b200.mnist import plan as P
    z, ch, nres, B, n_steps = _load("full")
    PG, BG, PD, PC = _params(ch, nres)
    ga = P.Arena(0, ch, nres, "cuda").load_dict([v.cuda() for v in PG.values()])
    da = P.Arena(1, ch, nres, "cuda").load_dict([v.cuda() for v in PD.values()])
    ca = P.Arena(2, ch, nres, "cuda").load_dict([v.cuda() for v in PC.values()])
    running = torch.zeros(2 * nres, 2, ch, device="cuda")
    running[:, 1] = 1
    nbt = torch.zeros(2 * nres, dtype=torch.int64, device="cuda")
    plan = P.MnistStepPlan(B, ga, da, ca, running, nbt, P.StepConfig(precision="bf16"), ch, nres)
    x, y, t, mask = (v.cuda().contiguous() for v in O.synth_batch(B, 700))

    def l2(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300)
    raw, masked = plan.g_forward(x, t, mask, True)
    assert l2(raw.cpu().numpy(), z["fwd_raw"]) < 1e-2 and l2(masked.cpu().numpy(), z["fwd_masked"]) < 1e-2
    assert l2(plan.d_forward(x, y).cpu().numpy(), z["fwd_d_logits"]) < 2e-2
    assert l2(plan.c_forward(x).cpu().numpy(), z["fwd_c_logits"]) < 2e-2
    ga.load_dict([v.cuda() for v in PG.values()])
    running.zero_()
    running[:, 1] = 1
    nbt.zero_()
    plan.refresh_weights()
    for i in range(n_steps):
        b = O.synth_batch(B, 800 + i, mnist_like=(i % 2 == 1))
        plan.step(*(v.cuda().contiguous() for v in b))
    torch.cuda.synchronize()
    gk, dk = list(O.g_param_shapes(ch, nres).items()), list(O.d_param_shapes().items())
    for key in z.files:
        if not key.startswith("train_"):
            continue
        net, k = key[6], key[8:]
        names = [n for n, _ in (gk if net == "G" else dk)]
        if k not in names or O.is_bn_shadowed_bias(k):
            continue
        i = names.index(k)
        got = _summary((ga if net == "G" else da).view(i, (gk if net == "G" else dk)[i][1]))
        ref, lr = z[key], (5e-5 if net == "G" else 1e-5)
        d = np.abs(got[2:] - ref[2:])
        assert d.max() <= 2.02 * lr * n_steps + 1e-7, (key, d.max())
        assert np.median(d) <= 0.5 * lr + 1e-7, (key, np.median(d))
        assert abs(got[1] - ref[1]) <= 1e-3 * abs(ref[1]) + 1e-6, key
