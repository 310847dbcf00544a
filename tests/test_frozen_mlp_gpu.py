"""csrc/frozen_mlp.cu: the frozen classifier's forward + cross-entropy + input gradient as one launch
(house_sales_kc_usa/trainer.py:300-303, moons/trainer.py:85-87) against float64 torch autograd."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def grad_close(got, ref):
    """A row whose pre-activation sits within rounding of a LeakyReLU kink takes the other branch in fp32 than in float64:
    all but a handful of rows (at most 0.2 %) agree to 1e-4 of the largest gradient; a flipped row's gradient is simply a
    different one, so no norm over all rows is asked for."""
    d = (got.double() - ref).abs().amax(1) / ref.abs().max()
    bad = (d > 1e-4).sum().item()
    assert bad <= max(1, got.shape[0] // 500), (bad, d.max().item())


@pytest.mark.parametrize("dims,slope", [([17, 256, 256, 128, 64, 4], 0.1), ([2, 32, 32, 3], 0.0), ([64, 128, 8], 0.2),
                                        ([5, 64, 256, 32, 128, 256, 2], 0.1), ([5, 32, 16, 16, 1], 0.2)])
@pytest.mark.parametrize("B", [4096, 1000, 64, 5])
def test_frozen_mlp_forward_loss_and_input_gradient(dims, slope, B):
    import pcg_b200  # noqa: F401
    from pcg_b200 import ops as K
    torch.manual_seed(B + sum(dims))
    dev = "cuda"
    Ws = [torch.randn(dims[j + 1], dims[j], device=dev) * (2.0 / dims[j]) ** 0.5 for j in range(len(dims) - 1)]
    bs = [torch.randn(dims[j + 1], device=dev) * 0.1 for j in range(len(dims) - 1)]
    x = torch.randn(B, dims[0], device=dev)
    target = torch.randint(0, dims[-1], (B,), device=dev)
    parts = K.frozen_mlp_parts(dims, B)
    assert parts == (B + 31) // 32 and K.frozen_mlp_parts([17, 96, 4], B) == -1 and K.frozen_mlp_parts([17, 64, 9], B) == -1
    loss_part = torch.full((parts,), 7.0, device=dev)
    dx, logits = torch.full((B, dims[0]), 7.0, device=dev), torch.full((B, dims[-1]), 7.0, device=dev)
    WTs = [w.t().contiguous() for w in Ws]
    K.frozen_mlp_ce_grad(Ws, WTs, bs, x, target, loss_part, dx, wgt=2.0, slope=slope, logits=logits)
    xd = x.double().requires_grad_(True)
    h = xd
    for j, (w, b) in enumerate(zip(Ws, bs)):
        h = h @ w.double().t() + b.double()
        if j + 1 < len(Ws):
            h = F.leaky_relu(h, slope)
    loss = F.cross_entropy(h, target)
    g, = torch.autograd.grad(2.0 * loss, xd, retain_graph=True)
    assert rel(logits, h) < 2e-5
    assert abs(loss_part.sum().item() / B - loss.item()) < 2e-5 * abs(loss.item()) + 1e-6
    grad_close(dx, g)
    K.frozen_mlp_ce_grad(Ws, WTs, bs, x, target, loss_part, dx, wgt=2.0, slope=slope)     # logits not stored
    grad_close(dx, g)
    # the mean of the outputs as the loss (a critic score in the generator step)
    g2, = torch.autograd.grad(-h.sum(1).mean(), xd)
    K.frozen_mlp_ce_grad(Ws, WTs, bs, x, None, loss_part, dx, wgt=-1.0, slope=slope, mean_output=True)
    grad_close(dx, g2)
    assert abs(loss_part.sum().item() / B - h.sum(1).mean().item()) < 2e-5 * abs(h.sum(1).mean().item()) + 1e-5
