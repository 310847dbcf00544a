"""Small fully connected layers (csrc/linear_small.cu) through pcg_conv_fprop / pcg_conv_dgrad (1x1 geometry) against
float64 torch: forward with bias + activation, data gradient with skip add and activation derivative, ragged row counts,
every column-count bucket (1..128); row counts above 2048 take the generic kernel (same contract, same test).  Layers of the tabular generators / critics: house_sales_kc_usa/models/generator.py:13-92,
discriminator.py:9-20, moons/models/*.py."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


@pytest.mark.parametrize("M,K,N", [(4096, 32, 32), (4096, 38, 32), (4096, 21, 32), (4096, 64, 128), (4096, 128, 64),
                                   (4096, 32, 2), (64, 7, 32), (1000, 32, 30), (37, 16, 2), (4096, 128, 1), (5, 1, 128),
                                   (777, 100, 70)])
def test_linear_small_forward_and_data_gradient(M, K, N):
    import pcg_b200  # noqa: F401
    from pcg_b200 import ops as Kk
    torch.manual_seed(M + K + N)
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * K ** -0.5
    b = torch.randn(N, device="cuda") * 0.1
    for act, fn in ((Kk.ACT_NONE, lambda t: t), (Kk.ACT_LRELU, lambda t: F.leaky_relu(t, 0.1)), (Kk.ACT_RELU, F.relu)):
        out = torch.full((M, N), 9.0, device="cuda")
        Kk.linear_fwd(x, w, out, b, act, 0.1)
        assert rel(out, fn(x.double() @ w.double().t() + b.double())) < 2e-5, (act,)
    out = torch.full((M, N), 9.0, device="cuda")
    Kk.linear_fwd(x, w, out)                                           # no bias
    assert rel(out, x.double() @ w.double().t()) < 2e-5
    # data gradient: dx = (dy @ W + skip) * lrelu'(ref)
    dy = torch.randn(M, N, device="cuda")
    wT = w.t().contiguous()                                            # [K][N], what pack_weights(..., wd=) produces
    skip, ref = torch.randn(M, K, device="cuda"), torch.randn(M, K, device="cuda")
    dx = torch.full((M, K), 9.0, device="cuda")
    Kk.linear_dgrad(dy, wT, dx, K)
    assert rel(dx, dy.double() @ w.double()) < 2e-5
    Kk.linear_dgrad(dy, wT, dx, K, act_ref=ref, ref_act=Kk.ACT_LRELU, ref_slope=0.2, add_src=skip)
    want = (dy.double() @ w.double() + skip.double()) * torch.where(ref.double() > 0, 1.0, 0.2)
    assert rel(dx, want) < 2e-5


@pytest.mark.parametrize("N,Cin,Cout,k", [(256, 512, 100, 4), (33, 24, 7, 4), (5, 130, 100, 3)])
def test_full_window_data_gradient_is_a_column_tiled_product(N, Cin, Cout, k):
    """ConvTranspose2d(100, 512, 4, 1, 0) on a 1x1 latent (mnist_dcgan.py:77) is the data gradient of a full-window
    convolution: din[n][(tap, ci)] = sum_co dout[n][co] * w[co][ci][tap].  Column tiles of 128 with permuted weight rows."""
    import pcg_b200  # noqa: F401
    from pcg_b200 import ops as Kk
    torch.manual_seed(N + Cin)
    w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.1                       # Conv2d(Cin -> Cout) layout
    wf = torch.empty(Cout * k * k * Cin, device="cuda")
    wd = torch.empty(Cin * k * k * Cout, device="cuda")
    Kk.pack_weights(w, k, wf=wf, wd=wd)
    dout = torch.randn(N, Cout, device="cuda")                                  # [N,1,1,Cout]
    din = torch.full((N, k, k, Cin), 9.0, device="cuda")
    skip, ref = torch.randn_like(din), torch.randn_like(din)
    Kk.conv_dgrad(dout, N, k, k, Cin, wd, Cout, k, 1, 0, din)
    want = torch.einsum("no,ocyx->nyxc", dout.double(), w.double())
    assert rel(din, want) < 2e-5
    Kk.conv_dgrad(dout, N, k, k, Cin, wd, Cout, k, 1, 0, din, add_src=skip, act_ref=ref, ref_act=Kk.ACT_RELU)
    assert rel(din, (want + skip.double()) * (ref.double() > 0)) < 2e-5


@pytest.mark.parametrize("M,K,N", [(4096, 32, 32), (4096, 38, 32), (4096, 21, 32), (4096, 64, 128), (4096, 128, 64),
                                   (4096, 128, 128), (4096, 128, 1), (64, 2, 32), (1000, 32, 7), (37, 17, 3), (5, 1, 1)])
def test_one_launch_weight_and_bias_gradient(M, K, N):
    """pcg_linear_wgrad_small against float64 torch; repeated calls reuse the scratch and are bit-identical (fixed
    summation order)."""
    import pcg_b200  # noqa: F401
    from pcg_b200 import ops as Kk
    torch.manual_seed(M + K + N)
    x, dy = torch.randn(M, K, device="cuda"), torch.randn(M, N, device="cuda")
    need = Kk.linear_wgrad_small_scratch_floats(M, K, N)
    assert need > 0 and Kk.linear_wgrad_small_scratch_floats(M, 129, N) == -1
    scratch = torch.full((need,), 123.0, device="cuda")                # no initialisation needed, may hold anything
    dw, db = torch.full((N, K), 9.0, device="cuda"), torch.full((N,), 9.0, device="cuda")
    Kk.linear_wgrad_small(x, dy, scratch, dw, db)
    assert rel(dw, dy.double().t() @ x.double()) < 2e-5 and rel(db, dy.double().sum(0)) < 2e-5
    dw2, db2 = torch.empty_like(dw), torch.empty_like(db)
    Kk.linear_wgrad_small(x, dy, scratch, dw2, db2)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)
    Kk.linear_wgrad_small(x, dy, scratch, dw2)                          # no bias gradient
    assert torch.equal(dw, dw2)
    # the dispatcher of ops.linear_wgrad takes it for layers of up to 4096 weights when the caller's scratch is large
    # enough; larger layers keep the primitive operators (same result to rounding)
    big = torch.zeros(max(need, Kk.conv_wgrad_scratch_floats(M, 1, 1, K, N, 1, 1, 0)) + 1024, device="cuda")
    dw3, db3 = torch.empty_like(dw), torch.empty_like(db)
    Kk.linear_wgrad(x, dy, big, dw3, db3, Kk.stat_scratch(max(N, 4), "cuda"))
    if N * K <= 4096:
        assert torch.equal(dw, dw3) and torch.equal(db, db3)
    else:
        assert rel(dw3, dw) < 2e-5 and rel(db3, db) < 2e-5
