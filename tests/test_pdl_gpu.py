"""Programmatic dependent launch (include/pcg.h pcg_set_pdl): every kernel of the library starts with
griddepcontrol.launch_dependents + griddepcontrol.wait, so launching with the programmatic-stream-serialization attribute
must not change any result.  The same operator-composed iteration (a ~100-kernel dependent chain) is run with the switch
off and on, eagerly and as a captured graph, and must agree bit for bit; likewise one tcgen05 convolution."""
import ctypes

import pytest
import torch

from oracle import moons_gan as M

pytestmark = pytest.mark.gpu


def _run(pdl, graph):
    import pcg_b200  # noqa: F401
    from pcg_b200 import _lib
    from pcg_b200.moons.gan import MlpGanPlan
    L = _lib.load()
    prev = L.pcg_set_pdl(1 if pdl else 0)
    try:
        gs, ds = M.shapes(label_dim=2)
        plan = MlpGanPlan(256, 32, 2, 128, "cuda", use_graph=graph, fused=False)
        plan.G.load({"net." + k: v for k, v in M.synth_params(gs, 1).items()})
        plan.D.load({"net." + k: v for k, v in M.synth_params(ds, 2).items()})
        plan.refresh()
        out = []
        for step in range(3):
            b = M.synth_batch(256, 300 + step, label_dim=2)
            out.append(plan.step(*[None if t is None else t.cuda().contiguous() for t in b]).clone())
        torch.cuda.synchronize()
        return torch.stack(out).cpu(), plan.G.data.clone().cpu(), plan.D.data.clone().cpu()
    finally:
        L.pcg_set_pdl(prev)


@pytest.mark.parametrize("graph", [False, True])
def test_pdl_does_not_change_results(graph):
    a = _run(False, graph)
    b = _run(True, graph)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_pdl_tensor_core_conv():
    import pcg_b200  # noqa: F401
    from pcg_b200 import _lib
    L, P, st = _lib.load(), _lib.ptr, _lib.stream_ptr()
    torch.manual_seed(0)
    N, HW = 8, 28
    x = torch.randn(N, HW, HW, 64, device="cuda").to(torch.bfloat16)
    w = torch.randn(64, 64, 3, 3, device="cuda") * 0.05
    wf = torch.empty(64, 9, 64, dtype=torch.bfloat16, device="cuda")
    wd = torch.empty(64, 9, 64, dtype=torch.bfloat16, device="cuda")
    outs = []
    for pdl in (0, 1):
        prev = L.pcg_set_pdl(pdl)
        try:
            _lib.check(L.pcg_pack_conv_weights_tc(P(w), 64, 64, 3, P(wf), P(wd), st))
            out = torch.zeros(N, HW, HW, 64, dtype=torch.bfloat16, device="cuda")
            for _ in range(3):          # back-to-back launches: the successor may start while the predecessor drains
                _lib.check(L.pcg_conv_tc64_fprop(P(x), N, HW, HW, P(wf), None, 0, ctypes.c_float(0.2), None, None, 0, P(out),
                                                 None, st))
            torch.cuda.synchronize()
            outs.append(out.float().cpu())
        finally:
            L.pcg_set_pdl(prev)
    assert torch.equal(outs[0], outs[1])
