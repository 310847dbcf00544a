"""Tensor-core operand cache (csrc/conv_auto.cu, pcg_set_operand_cache): bf16 conversions of unchanged operands are
reused, a write through any operator of pcg_b200.ops invalidates them, producers (BatchNorm apply) fill them directly, and
a clear drops everything (tensors written by torch)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def test_reuse_invalidation_and_producer_fill():
    import pcg_b200  # noqa: F401
    from pcg_b200 import _lib, ops as K
    torch.manual_seed(0)
    N, H, Ci, Co = 8, 8, 64, 64
    x = torch.randn(N, H, H, Ci, device="cuda")
    w = torch.randn(Co, Ci, 3, 3, device="cuda") * 0.05
    wf = torch.empty(Co * 9 * Ci, device="cuda")
    K.pack_weights(w, 3, wf=wf)
    out = torch.empty(N, H, H, Co, device="cuda")

    def ref(t):
        return F.conv2d(t.permute(0, 3, 1, 2).bfloat16().double(), w.bfloat16().double(), None, 1, 1).permute(0, 2, 3, 1)

    K.set_conv_tensor_cores(True)
    prev_terms = K.set_conv_tensor_core_terms(1)
    prev = K.set_operand_cache(True)
    try:
        K.operand_cache_clear()
        n0 = _lib.launch_count()
        K.conv_fprop(x, N, H, H, Ci, wf, Co, 3, 1, 1, out)
        first = _lib.launch_count() - n0                     # two conversions + the convolution
        assert rel(out, ref(x)) < 2e-3
        n0 = _lib.launch_count()
        K.conv_fprop(x, N, H, H, Ci, wf, Co, 3, 1, 1, out)
        assert _lib.launch_count() - n0 == first - 2         # both operands reused
        # an operator of this module writes x: its conversion is dropped, the weight's is kept
        K.unary(x, K.SCALE, x, 2.0)
        n0 = _lib.launch_count()
        K.conv_fprop(x, N, H, H, Ci, wf, Co, 3, 1, 1, out)
        assert _lib.launch_count() - n0 == first - 1
        assert rel(out, ref(x)) < 2e-3
        # torch writes x behind the library's back: stale until the caller clears (what every plan body does first)
        x.mul_(0.5)
        K.conv_fprop(x, N, H, H, Ci, wf, Co, 3, 1, 1, out)
        assert rel(out, ref(x * 2.0)) < 2e-3                 # still the old operand
        K.operand_cache_clear()
        K.conv_fprop(x, N, H, H, Ci, wf, Co, 3, 1, 1, out)
        assert rel(out, ref(x)) < 2e-3
        # a producer fills the conversion a consumer asked for: BatchNorm apply writes z and its bf16 copy
        y = torch.randn(N, H, H, Ci, device="cuda")
        z = torch.empty_like(y)
        gam, bet = torch.ones(Ci, device="cuda"), torch.zeros(Ci, device="cuda")
        rm, rv, nbt = torch.zeros(Ci, device="cuda"), torch.ones(Ci, device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
        st = K.BNState(Ci, "cuda")
        K.bn_train_fwd(y, N * H * H, Ci, gam, bet, rm, rv, nbt, st, z, act=K.ACT_RELU)
        K.conv_fprop(z, N, H, H, Ci, wf, Co, 3, 1, 1, out)   # registers (and converts) z
        y.normal_()
        K.operand_cache_clear()
        K.conv_fprop(x, N, H, H, Ci, wf, Co, 3, 1, 1, out)   # weight conversion valid again
        K.bn_train_fwd(y, N * H * H, Ci, gam, bet, rm, rv, nbt, st, z, act=K.ACT_RELU)
        n0 = _lib.launch_count()
        K.conv_fprop(z, N, H, H, Ci, wf, Co, 3, 1, 1, out)
        assert _lib.launch_count() - n0 == first - 2         # no conversion pass for z: the apply kernel wrote it
        assert rel(out, ref(z)) < 2e-3
    finally:
        K.set_conv_tensor_cores(False)
        K.set_conv_tensor_core_terms(prev_terms)
        K.set_operand_cache(prev)
    # off again: every call converts
    K.set_conv_tensor_cores(True)
    K.set_conv_tensor_core_terms(1)
    try:
        n0 = _lib.launch_count()
        K.conv_fprop(x, N, H, H, Ci, wf, Co, 3, 1, 1, out)
        assert _lib.launch_count() - n0 == first
    finally:
        K.set_conv_tensor_cores(False)
        K.set_conv_tensor_core_terms(prev_terms)
