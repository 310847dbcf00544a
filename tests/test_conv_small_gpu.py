"""Direct parity of the skinny-layer kernels (csrc/conv_small.cu) through their C-ABI test entry points against a
plain PyTorch fp32 reference of the same op on the same bf16-rounded operands (so only the accumulation order differs:
tolerance 2e-3 relative to the tensor's max).  Ragged batch sizes exercise tiles that run past the tensor end."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
_f = ctypes.c_float


def _env():
    import pcg_b200  # noqa: F401
    from pcg_b200 import _lib
    return _lib.load(), _lib.ptr, _lib.stream_ptr(), _lib.check


def rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).abs().max() / (b.float().abs().max() + 1e-30)).item()


def bf(t):
    return t.to(torch.bfloat16)


def nhwc(x):      # NCHW -> NHWC contiguous
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("N,Cin", [(3, 64), (5, 32), (16, 64)])
def test_conv_to1(N, Cin):
    L, P, st, chk = _env()
    torch.manual_seed(N)
    x = bf(torch.randn(N, Cin, 28, 28, device="cuda"))
    w = bf(torch.randn(1, Cin, 3, 3, device="cuda") * 0.1)
    b = torch.randn(1, device="cuda")
    w9 = w[0].permute(1, 2, 0).reshape(9, Cin).contiguous()            # [tap][c]
    out = torch.empty(N, 28, 28, device="cuda")
    chk(L.pcg_conv_to1(P(nhwc(x)), N, 28, 28, Cin, P(w9), P(b), P(out), st))
    ref = F.conv2d(x.float(), w.float(), b, padding=1)[:, 0]
    assert rel(out, ref) < 2e-3


@pytest.mark.parametrize("N,Cs,Cout,stride,f32", [(3, 3, 64, 1, False), (5, 2, 64, 2, False), (4, 1, 64, 1, False),
                                                  (3, 1, 32, 1, True)])
def test_conv_few(N, Cs, Cout, stride, f32):
    L, P, st, chk = _env()
    torch.manual_seed(Cs * 10 + N)
    x = torch.randn(N, Cs, 28, 28, device="cuda")
    x = x if f32 else bf(x)
    w = bf(torch.randn(Cout, Cs, 3, 3, device="cuda") * 0.2)
    b = torch.randn(Cout, device="cuda")
    wnk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cs).contiguous()     # [co][tap*Cs + c]
    Ho = (28 + 2 - 3) // stride + 1
    ref_pre = F.conv2d(bf(x).float(), w.float(), b, stride=stride, padding=1)
    ref = F.leaky_relu(ref_pre, 0.2)
    act_ref = bf(torch.randn(N, Ho, Ho, Cout, device="cuda"))
    ref = ref * torch.where(act_ref.float() > 0, 1.0, 0.2).permute(0, 3, 1, 2)
    out = torch.empty(N, Ho, Ho, Cout, device="cuda", dtype=torch.bfloat16)
    chk(L.pcg_conv_few(P(nhwc(x)), 1 if f32 else 0, N, 28, 28, Cs, P(wnk), Cout, stride, P(b), 1, _f(0.2), P(act_ref), 1,
                       _f(0.2), P(out), st))
    assert rel(out.permute(0, 3, 1, 2), ref) < 1e-2                    # bf16 output rounding (2^-9) dominates


@pytest.mark.parametrize("N", [3, 9])
def test_dgrad_s2_to1(N):
    L, P, st, chk = _env()
    torch.manual_seed(N)
    Cin, Cout = 2, 64
    w = bf(torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.1)
    dy = bf(torch.randn(N, Cout, 14, 14, device="cuda"))
    ref = torch.nn.grad.conv2d_input((N, Cin, 28, 28), w.float(), dy.float(), stride=2, padding=1)
    for ch in range(Cin):
        # rotated packing row of channel ch: [8 - tap][co]
        wrot = w[:, ch].reshape(Cout, 9).t().flip(0).contiguous()
        dx = torch.empty(N, 28, 28, device="cuda")
        chk(L.pcg_dgrad_s2_to1(P(nhwc(dy)), N, 28, 28, P(wrot), P(dx), st))
        assert rel(dx, ref[:, ch]) < 2e-3, ch


@pytest.mark.parametrize("N,Cs,stride", [(3, 3, 1), (5, 2, 2)])
def test_wgrad_few(N, Cs, stride):
    L, P, st, chk = _env()
    torch.manual_seed(N + Cs)
    x = bf(torch.randn(N, Cs, 28, 28, device="cuda"))
    Ho = (28 + 2 - 3) // stride + 1
    dy = bf(torch.randn(N, 64, Ho, Ho, device="cuda"))
    ref = torch.nn.grad.conv2d_weight(x.float(), (64, Cs, 3, 3), dy.float(), stride=stride, padding=1)
    part = torch.zeros(L.pcg_wgrad_small_parts() * 2048, device="cuda")
    dw, db = torch.empty(64, Cs, 3, 3, device="cuda"), torch.empty(64, device="cuda")
    chk(L.pcg_wgrad_few(P(nhwc(x)), P(nhwc(dy)), N, 28, 28, Cs, stride, P(part), P(dw), P(db), st))
    assert rel(dw, ref) < 2e-3 and rel(db, dy.float().sum(dim=(0, 2, 3))) < 2e-3


@pytest.mark.parametrize("N", [3, 7])
def test_wgrad_to1(N):
    L, P, st, chk = _env()
    torch.manual_seed(N)
    x = bf(torch.randn(N, 64, 28, 28, device="cuda"))
    g = bf(torch.randn(N, 1, 28, 28, device="cuda"))
    ref = torch.nn.grad.conv2d_weight(x.float(), (1, 64, 3, 3), g.float(), padding=1)
    part = torch.zeros(L.pcg_wgrad_small_parts() * 2048, device="cuda")
    dw, db = torch.empty(1, 64, 3, 3, device="cuda"), torch.empty(1, device="cuda")
    chk(L.pcg_wgrad_to1(P(nhwc(x)), P(g.reshape(N, 28, 28).contiguous()), N, 28, 28, P(part), P(dw), P(db), st))
    assert rel(dw, ref) < 2e-3 and abs(db.item() - g.float().sum().item()) < 2e-3 * g.float().abs().sum().item()
