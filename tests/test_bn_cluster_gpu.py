"""One-launch cluster BatchNorm (csrc/bn_cluster.cu) through pcg_bn_train_fwd / pcg_bn_train_bwd against
torch.nn.functional.batch_norm + autograd in float64 (torch.nn.BatchNorm1d training mode, generator.py:29-35 of the KC
experiment and moons/models/generator.py:9-17): output, saved statistics, running buffers, num_batches_tracked, the
input gradient and the affine gradients.  Shapes: the moons generator (64 x 32, 64 x 16) and a ragged small one (cluster
path), the KC generator (4096 x 32) and larger ones (three-launch pipeline: same contract, same test)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


@pytest.mark.parametrize("M,C,act", [(4096, 32, 0), (4096, 32, 2), (64, 32, 2), (64, 16, 1), (1000, 64, 0), (37, 8, 2),
                                     (20000, 32, 2)])
def test_bn_train_fwd_bwd_match_torch(M, C, act):
    import pcg_b200  # noqa: F401
    from pcg_b200 import ops as K
    torch.manual_seed(M + C + act)
    y = (torch.randn(M, C, device="cuda") * 1.7 + 0.4).contiguous()
    gamma, beta = torch.randn(C, device="cuda"), torch.randn(C, device="cuda") * 0.3
    rm, rv = torch.randn(C, device="cuda") * 0.1, torch.rand(C, device="cuda") + 0.5
    nbt = torch.tensor(3, dtype=torch.int64, device="cuda")
    st = K.BNState(C, "cuda")
    z = torch.full((M, C), 7.0, device="cuda")
    rm0, rv0 = rm.clone(), rv.clone()
    K.bn_train_fwd(y, M, C, gamma, beta, rm, rv, nbt, st, z, act=act, slope=0.2)
    torch.cuda.synchronize()
    yd = y.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rmd, rvd = rm0.double().clone(), rv0.double().clone()
    ref = F.batch_norm(yd, rmd, rvd, gd, bd, training=True, momentum=0.1, eps=1e-5)
    ref_a = F.leaky_relu(ref, 0.2) if act == 1 else F.relu(ref) if act == 2 else ref
    assert rel(z, ref_a) < 2e-5
    assert rel(rm, rmd) < 1e-5 and rel(rv, rvd) < 1e-5 and int(nbt) == 4
    assert rel(st.mean, yd.mean(0)) < 1e-5
    assert rel(st.rstd, (yd.var(0, unbiased=False) + 1e-5).rsqrt()) < 1e-5
    dz = torch.randn(M, C, device="cuda")
    dy = torch.full((M, C), 7.0, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    K.bn_train_bwd(dz, y, M, C, gamma, st, dy, dg, db, gscale=0.5, act=act, slope=0.2)
    torch.cuda.synchronize()
    gy, gg, gb = torch.autograd.grad(ref_a, (yd, gd, bd), dz.double() * 0.5)
    assert rel(dy, gy) < 5e-5 and rel(dg, gg) < 5e-5 and rel(db, gb) < 5e-5
