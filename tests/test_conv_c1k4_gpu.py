"""One-channel 4x4 / stride-2 kernels (csrc/conv_c1k4.cu) through the primitive-operator C ABI (pcg_conv_fprop /
pcg_conv_dgrad / pcg_conv_wgrad pick them for the geometry [N,H,W,1] <-> [N,H/2,W/2,64]) against plain PyTorch fp32:
DCGAN's Conv2d(1, 64, 4, 2, 1) and ConvTranspose2d(64, 1, 4, 2, 1), dconv_gan/mnist/mnist_dcgan.py:89,100.
The reference runs in float64 (torch's own fp32 convolutions use TF32 on this GPU: 3e-4); tolerance 2e-5 relative to the
tensor's max."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


@pytest.mark.parametrize("N,HW", [(3, 64), (256, 64), (5, 16), (2, 32)])
def test_c1k4_operators_match_torch(N, HW):
    import pcg_b200  # noqa: F401
    from pcg_b200 import ops as K
    torch.manual_seed(N + HW)
    w = torch.randn(64, 1, 4, 4, device="cuda") * 0.2
    x = torch.randn(N, 1, HW, HW, device="cuda")
    w64, x64 = w.double(), x.double()
    Ho = HW // 2
    wf = torch.empty(64 * 16, device="cuda")
    wd = torch.empty(64 * 16, device="cuda")
    K.pack_weights(w, 4, wf=wf, wd=wd)
    xn = x.permute(0, 2, 3, 1).contiguous()                       # NHWC (C = 1: same bytes)
    # forward (+ LeakyReLU), as D's first layer
    out = torch.full((N, Ho, Ho, 64), 9.0, device="cuda")
    K.conv_fprop(xn, N, HW, HW, 1, wf, 64, 4, 2, 1, out, act=K.ACT_LRELU, slope=0.2)
    ref = F.leaky_relu(F.conv2d(x64, w64, None, 2, 1), 0.2)
    assert rel(out.permute(0, 3, 1, 2), ref) < 2e-5
    # plain forward, as the input gradient of G's last ConvTranspose2d
    K.conv_fprop(xn, N, HW, HW, 1, wf, 64, 4, 2, 1, out)
    assert rel(out.permute(0, 3, 1, 2), F.conv2d(x64, w64, None, 2, 1)) < 2e-5
    # data gradient = ConvTranspose2d(64, 1, 4, 2, 1) forward with the same weight tensor
    dy = torch.randn(N, 64, Ho, Ho, device="cuda")
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    dx = torch.full((N, HW, HW, 1), 9.0, device="cuda")
    K.conv_dgrad(dyn, N, HW, HW, 1, wd, 64, 4, 2, 1, dx)
    assert rel(dx.view(N, 1, HW, HW), F.conv_transpose2d(dy.double(), w64, None, 2, 1)) < 2e-5
    # weight gradient
    scratch = K.conv_wgrad_scratch(N, HW, HW, 1, 64, 4, 2, 1, "cuda")
    dw = torch.full((64, 1, 4, 4), 9.0, device="cuda")
    K.conv_wgrad(xn, dyn, N, HW, HW, 1, 64, 4, 2, 1, scratch, dw)
    wz = w64.clone().requires_grad_(True)
    (gw,) = torch.autograd.grad(F.conv2d(x64, wz, None, 2, 1), wz, dy.double())
    assert rel(dw, gw) < 5e-5


@pytest.mark.parametrize("N,C,k", [(256, 512, 4), (7, 64, 4), (3, 8, 2)])
def test_full_window_conv_to_one_output(N, C, k):
    """Conv2d(C, 1, k, 1, 0) on a k x k map (DCGAN's last discriminator layer, mnist_dcgan.py:112): forward, data
    gradient, weight gradient through pcg_conv_fprop / dgrad / wgrad against float64 torch."""
    import pcg_b200  # noqa: F401
    from pcg_b200 import ops as K
    torch.manual_seed(N + C)
    w = torch.randn(1, C, k, k, device="cuda") * 0.05
    x = torch.randn(N, C, k, k, device="cuda")
    wf, wd = torch.empty(C * k * k, device="cuda"), torch.empty(C * k * k, device="cuda")
    K.pack_weights(w, k, wf=wf, wd=wd)
    xn = x.permute(0, 2, 3, 1).contiguous()
    out = torch.full((N,), 9.0, device="cuda")
    K.conv_fprop(xn, N, k, k, C, wf, 1, k, 1, 0, out)
    assert rel(out, F.conv2d(x.double(), w.double()).view(N)) < 2e-5
    dz = torch.randn(N, device="cuda")
    dx = torch.full((N, k, k, C), 9.0, device="cuda")
    K.conv_dgrad(dz, N, k, k, C, wd, 1, k, 1, 0, dx)
    ref = F.conv_transpose2d(dz.double().view(N, 1, 1, 1), w.double())
    assert rel(dx.permute(0, 3, 1, 2), ref) < 2e-5
    scratch = K.conv_wgrad_scratch(N, k, k, C, 1, k, 1, 0, "cuda")
    dw = torch.full((1, C, k, k), 9.0, device="cuda")
    K.conv_wgrad(xn, dz, N, k, k, C, 1, k, 1, 0, scratch, dw)
    wz = w.double().clone().requires_grad_(True)
    (gw,) = torch.autograd.grad(F.conv2d(x.double(), wz), wz, dz.double().view(N, 1, 1, 1))
    assert rel(dw, gw) < 2e-5
