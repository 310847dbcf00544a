"""Direct parity of the 64->64 halo-tile tcgen05 kernels (csrc/conv_tc64.cu) through their C-ABI entry points against a
plain PyTorch fp32 reference of the same op on the same bf16-rounded operands (only the accumulation order and the final
bf16 rounding differ: tolerance 1e-2 relative to the tensor's max, 2e-3 on the fp32 statistics / weight gradients).

Both forward kernels are covered: the row-class stacked one (default when H % 4 == 0) and the one-class-per-tile one
(variant bit 256, and every other height).  Replaces nn.Conv2d(64, 64, 3, padding=1) forward and
ConvolutionBackward0 of conditional_counteRGAN/mnist/models/generator.py:11,14,49.  Ragged batch sizes exercise
super-tiles whose last rows fall outside the image and grids smaller / larger than the SM count."""
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
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-30)).item()


def nhwc_bf16(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(x):
    return x.float().permute(0, 3, 1, 2)


def pack(L, P, st, check, w):
    Cout, Cin, k, _ = w.shape
    f = torch.empty(Cout, k * k, Cin, dtype=torch.bfloat16, device="cuda")
    d = torch.empty(Cin, k * k, Cout, dtype=torch.bfloat16, device="cuda")
    check(L.pcg_pack_conv_weights_tc(P(w), Cout, Cin, k, P(f), P(d), st))
    return f, d


@pytest.mark.parametrize("variant", [0, 256])
@pytest.mark.parametrize("N,HW", [(3, 28), (8, 28), (301, 28), (5, 12), (5, 14)])
def test_fprop_bias_stats_and_epilogues(N, HW, variant):
    L, P, st, check = _env()
    L.pcg_conv_tc64_set_variant(variant)
    try:
        torch.manual_seed(N * 100 + HW)
        x = torch.randn(N, 64, HW, HW, device="cuda")
        w = torch.randn(64, 64, 3, 3, device="cuda") * (2.0 / 576) ** 0.5
        b = torch.randn(64, device="cuda") * 0.1
        xn = nhwc_bf16(x)
        wf, wd = pack(L, P, st, check, w)
        wb = w.to(torch.bfloat16).float()
        rows = L.pcg_conv_tc64_fprop_grid(N, HW, HW)
        assert rows >= 1
        # forward + bias + BatchNorm statistics
        out = torch.full((N, HW, HW, 64), 7.0, dtype=torch.bfloat16, device="cuda")
        stats = torch.full((rows, 128), 1e9, device="cuda")
        check(L.pcg_conv_tc64_fprop(P(xn), N, HW, HW, P(wf), P(b), 0, _f(0.2), None, None, 0, P(out), P(stats), st))
        torch.cuda.synchronize()
        ref = F.conv2d(nchw(xn), wb, b, padding=1)
        assert rel(nchw(out), ref) < 1e-2
        s = stats.double().sum(0)
        assert rel(s[:64], ref.double().sum(dim=(0, 2, 3))) < 2e-3
        assert rel(s[64:], (ref.double() ** 2).sum(dim=(0, 2, 3))) < 2e-3
        # forward + bias + LeakyReLU + residual add (eval-mode residual block epilogue)
        res = nhwc_bf16(torch.randn(N, 64, HW, HW, device="cuda"))
        out2 = torch.empty_like(out)
        check(L.pcg_conv_tc64_fprop(P(xn), N, HW, HW, P(wf), P(b), 1, _f(0.2), P(res), None, 0, P(out2), None, st))
        torch.cuda.synchronize()
        ref2 = F.leaky_relu(ref, 0.2) + nchw(res)
        assert rel(nchw(out2), ref2) < 1e-2
        # data gradient (rotated packing) times the LeakyReLU derivative taken from an activation reference
        dy = nhwc_bf16(torch.randn(N, 64, HW, HW, device="cuda") * 0.1)
        aref = nhwc_bf16(torch.randn(N, 64, HW, HW, device="cuda"))
        dx = torch.empty_like(out)
        check(L.pcg_conv_tc64_fprop(P(dy), N, HW, HW, P(wd), None, 0, _f(0.2), None, P(aref), 1, P(dx), None, st))
        torch.cuda.synchronize()
        gref = F.conv_transpose2d(nchw(dy), wb, None, padding=1)
        gref = gref * torch.where(nchw(aref) > 0, 1.0, 0.2)
        assert rel(nchw(dx), gref) < 1e-2
        # data gradient + skip-connection gradient
        dx2 = torch.empty_like(out)
        check(L.pcg_conv_tc64_fprop(P(dy), N, HW, HW, P(wd), None, 0, _f(0.2), P(res), None, 0, P(dx2), None, st))
        torch.cuda.synchronize()
        assert rel(nchw(dx2), F.conv_transpose2d(nchw(dy), wb, None, padding=1) + nchw(res)) < 1e-2
    finally:
        L.pcg_conv_tc64_set_variant(0)


@pytest.mark.parametrize("with_add", [False, True])
@pytest.mark.parametrize("N,HW", [(3, 28), (8, 28), (301, 28), (5, 12)])
def test_dgrad_with_fused_batchnorm_backward_reduction(N, HW, with_add):
    """Data gradient whose epilogue also takes the two column sums of the BatchNorm backward that consumes it, with and
    without the skip-connection gradient added first (two epilogue operands: the in-place staging ring).  Reference:
    plain torch fp32 on the same bf16 operands; the sums are taken of the fp32 result before its bf16 rounding."""
    L, P, st, check = _env()
    torch.manual_seed(N * 7 + HW + int(with_add))
    dy = nhwc_bf16(torch.randn(N, 64, HW, HW, device="cuda") * 0.1)
    y = nhwc_bf16(torch.randn(N, 64, HW, HW, device="cuda") * 1.5 + 0.3)
    add = nhwc_bf16(torch.randn(N, 64, HW, HW, device="cuda") * 0.2) if with_add else None
    w = torch.randn(64, 64, 3, 3, device="cuda") * (2.0 / 576) ** 0.5
    _, wd = pack(L, P, st, check, w)
    wb = w.to(torch.bfloat16).float()
    mean, var = torch.randn(64, device="cuda") * 0.2, torch.rand(64, device="cuda") + 0.5
    rstd = (var + 1e-5).rsqrt()
    gamma, beta = torch.randn(64, device="cuda"), torch.randn(64, device="cuda") * 0.3
    scale, shift = gamma * rstd, beta - mean * gamma * rstd
    rows = L.pcg_conv_tc64_fprop_grid(N, HW, HW)
    for act, gscale in ((1, 1.0), (0, 0.1)):          # BN1 (LeakyReLU in front of it in the backward) / BN2 (0.1 * dh)
        out = torch.full((N, HW, HW, 64), 3.0, dtype=torch.bfloat16, device="cuda")
        stats = torch.full((rows, 128), 1e9, device="cuda")
        check(L.pcg_conv_tc64_dgrad_bnred(P(dy), N, HW, HW, P(wd), P(add), P(y), P(mean), P(rstd), P(scale), P(shift), act,
                                          _f(0.2), _f(gscale), P(out), P(stats), st))
        torch.cuda.synchronize()
        v = F.conv_transpose2d(nchw(dy), wb, None, padding=1)
        if with_add:
            v = v + nchw(add)
        assert rel(nchw(out), v) < 1e-2
        yn = nchw(y).double()
        pre = yn * scale.double().view(1, 64, 1, 1) + shift.double().view(1, 64, 1, 1)
        g = v.double() * gscale * (torch.where(pre > 0, 1.0, 0.2) if act == 1 else 1.0)
        xhat = (yn - mean.double().view(1, 64, 1, 1)) * rstd.double().view(1, 64, 1, 1)
        s = stats.double().sum(0)
        assert rel(s[:64], g.sum(dim=(0, 2, 3))) < 2e-3, (act, rel(s[:64], g.sum(dim=(0, 2, 3))))
        assert rel(s[64:], (g * xhat).sum(dim=(0, 2, 3))) < 2e-3, (act, rel(s[64:], (g * xhat).sum(dim=(0, 2, 3))))


def test_stacked_and_one_class_kernels_agree():
    """Same products, same fp32 accumulation per output element up to the order of the nine taps: the two kernels
    must agree to bf16 rounding of identical fp32 sums almost everywhere (<= 1 bf16 ulp)."""
    L, P, st, check = _env()
    torch.manual_seed(5)
    N, HW = 16, 28
    xn = nhwc_bf16(torch.randn(N, 64, HW, HW, device="cuda"))
    w = torch.randn(64, 64, 3, 3, device="cuda") * 0.05
    wf, _ = pack(L, P, st, check, w)
    outs = []
    for variant in (0, 256):
        L.pcg_conv_tc64_set_variant(variant)
        out = torch.empty(N, HW, HW, 64, dtype=torch.bfloat16, device="cuda")
        check(L.pcg_conv_tc64_fprop(P(xn), N, HW, HW, P(wf), None, 0, _f(0.2), None, None, 0, P(out), None, st))
        torch.cuda.synchronize()
        outs.append(out.float())
    L.pcg_conv_tc64_set_variant(0)
    assert rel(outs[0], outs[1]) < 1e-2
    assert (outs[0] != outs[1]).float().mean().item() < 0.05


@pytest.mark.parametrize("variant", [0, 1024, 2048])
@pytest.mark.parametrize("N,HW", [(3, 28), (8, 28), (301, 28), (5, 14), (5, 12)])
def test_wgrad(N, HW, variant):
    """variant 0: shifted views on both operands (16 MMAs per tile), partials staged in shared memory and sent off by
    bulk copies (chunk-swizzled rows); 2048: the same MMAs, partials by direct stores; 1024: two taps per accumulator."""
    L, P, st, check = _env()
    L.pcg_conv_tc64_set_variant(variant)
    torch.manual_seed(N)
    xn = nhwc_bf16(torch.randn(N, 64, HW, HW, device="cuda"))
    dyn = nhwc_bf16(torch.randn(N, 64, HW, HW, device="cuda") * 0.1)
    grid = L.pcg_conv_tc64_grid(N, HW, HW)
    part = torch.zeros(grid, 9 * 64 * 64, device="cuda")
    dw = torch.zeros(64, 64, 3, 3, device="cuda")
    check(L.pcg_conv_tc64_wgrad(P(xn), P(dyn), N, HW, HW, P(part), P(dw), st))
    torch.cuda.synchronize()
    wz = torch.zeros(64, 64, 3, 3, device="cuda", requires_grad=True)
    yy = F.conv2d(nchw(xn), wz, None, padding=1)
    (gw,) = torch.autograd.grad(yy, wz, nchw(dyn))
    L.pcg_conv_tc64_set_variant(0)
    assert rel(dw, gw) < 2e-3


@pytest.mark.parametrize("N,HW", [(8, 28), (301, 28), (5, 12)])
def test_wgrad_staged_writeout_is_bit_identical(N, HW):
    """The staged, chunk-swizzled write-out only changes where a partial sum is stored: dW must not change by a bit."""
    L, P, st, check = _env()
    torch.manual_seed(N + 1)
    xn = nhwc_bf16(torch.randn(N, 64, HW, HW, device="cuda"))
    dyn = nhwc_bf16(torch.randn(N, 64, HW, HW, device="cuda") * 0.1)
    grid = L.pcg_conv_tc64_grid(N, HW, HW)
    outs = []
    for variant in (0, 2048):
        L.pcg_conv_tc64_set_variant(variant)
        part = torch.full((grid, 9 * 64 * 64), float("nan"), device="cuda")
        dw = torch.zeros(64, 64, 3, 3, device="cuda")
        check(L.pcg_conv_tc64_wgrad(P(xn), P(dyn), N, HW, HW, P(part), P(dw), st))
        torch.cuda.synchronize()
        assert torch.isfinite(part).all()          # every element of every partial is written
        outs.append(dw)
    L.pcg_conv_tc64_set_variant(0)
    assert torch.equal(outs[0], outs[1])
