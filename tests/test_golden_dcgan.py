"""Golden vectors of the DCGAN iteration produced by the reference script's own classes and loop body
(tests/golden/make_golden_dcgan.py; batch 4, 2 iterations):
  * CPU (-m "not gpu"): the oracle reproduces them -> oracle/dcgan.py stays pinned where /root/reference is absent;
  * GPU (-m gpu): the native plan (exact fp32 CUDA-core mode and the tensor-core mode with bf16x3 operands) reproduces
    them directly.
Per tensor the fixture holds [sum, abs-sum, 32 samples]; samples are compared in units of the Adam step (lr = 2e-4: an
element with a rounding-level gradient may move by +-lr per step on either side)."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import dcgan as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dcgan.npz")
LR = 2e-4


def _summary(t):
    t = t.detach().double().flatten().cpu()
    idx = torch.linspace(0, t.numel() - 1, 32).long()
    return np.concatenate([[t.sum().item(), t.abs().sum().item()], t[idx].numpy()])


def _check(z, get, steps, med_tol, abs_tol, max_steps=2.02):
    for key in z.files:
        if key in ("meta", "errs"):
            continue
        net, k = key[0], key[2:]
        ref = z[key]
        if "num_batches" in k:
            continue
        got = _summary(get(net, k))
        if "running" in k:
            assert np.allclose(got[2:], ref[2:], rtol=5e-3, atol=1e-4), key
            continue
        d = np.abs(got[2:] - ref[2:])
        assert d.max() <= max_steps * LR * steps + 1e-7, (key, d.max())
        assert np.median(d) <= med_tol * LR + 1e-7, (key, np.median(d))
        assert abs(got[1] - ref[1]) <= abs_tol * abs(ref[1]) + 1e-6, key


def test_oracle_reproduces_reference_golden():
    z = np.load(GOLD)
    B, steps, seed0 = (int(v) for v in z["meta"])
    S = O.make_state(O.synth_params(O.g_shapes(), 5), O.buffers(O.g_shapes()), O.synth_params(O.d_shapes(), 6),
                     O.buffers(O.d_shapes()))
    for it in range(steps):
        sc, _ = O.dcgan_step(S, *O.synth_batch(B, seed0 + it))
        tol = 2e-5 if it == 0 else 2e-3     # later iterations sit behind Adam steps whose rounding-level elements may flip
        assert abs(sc["errD"] - z["errs"][it][0]) < tol * abs(z["errs"][it][0]) + 1e-6
        assert abs(sc["errG"] - z["errs"][it][1]) < tol * abs(z["errs"][it][1]) + 1e-6
    get = lambda net, k: (S[net][k] if k in S[net] else S[net + "B"][k])  # noqa: E731
    _check(z, get, steps, 0.02, 3e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("tc", [False, True])
def test_native_plan_reproduces_reference_golden(tc):
    import pcg_b200  # noqa: F401
    from pcg_b200.dcgan import DcganPlan
    z = np.load(GOLD)
    B, steps, seed0 = (int(v) for v in z["meta"])
    plan = DcganPlan(B, "cuda", use_graph=False, tensor_cores=tc, operand_terms=3)      # bf16x3: the fp32-level mode
    plan.G.load(O.synth_params(O.g_shapes(), 5))
    plan.D.load(O.synth_params(O.d_shapes(), 6))
    plan.refresh()
    for it in range(steps):
        real, noise = O.synth_batch(B, seed0 + it)
        got = plan.step(real.cuda(), noise.cuda()).tolist()
        tol = 2e-4 if it == 0 else 3e-2
        assert abs(got[0] - z["errs"][it][0]) <= tol * abs(z["errs"][it][0]) + 1e-6, (it, got[0], z["errs"][it][0])
        assert abs(got[1] - z["errs"][it][1]) <= tol * abs(z["errs"][it][1]) + 1e-6, (it, got[1], z["errs"][it][1])
    torch.cuda.synchronize()
    flats = {"G": plan.G, "D": plan.D}
    sub = OrderedDict((key, None) for key in z.files if key not in ("meta", "errs") and key[2:] in flats[key[0]].shapes)
    assert len(sub) >= 20

    class _Z:
        files = list(sub)

        def __getitem__(self, k):
            return z[k]
    # with beta1 = 0.5 a bias-corrected Adam step can exceed lr by ~15 %, and the two sides may step in opposite directions
    _check(_Z(), lambda net, k: flats[net].p(k), steps, 0.1, 2e-3, max_steps=2.5)
