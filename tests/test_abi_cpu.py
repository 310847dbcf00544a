"""CPU-side checks: libpcg.so loads and exports every symbol include/pcg.h declares; layouts match the
reference's parameter counts; the mirror modules keep the reference's state_dict keys and init."""
import ctypes
import os
import re
from collections import OrderedDict

import pytest
import torch

import pcg_b200  # noqa: F401
from pcg_b200 import _lib
from pcg_b200.mnist import plan as P
from oracle import mnist_countergan as O
from tests._refload import experiment, have_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pcg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/pcg.h but not exported by libpcg.so"
    assert L.pcg_version() == 1


def test_layouts_match_reference_parameter_counts():
    for net, shapes, count in ((0, O.g_param_shapes(), 491809), (1, O.d_param_shapes(), 967713),
                               (2, O.c_param_shapes(), 1701130)):
        slots, total = P.layout(net)
        assert len(slots) == len(shapes)
        assert sum(n for _, n in slots) == count
        for (off, n), shp in zip(slots, shapes.values()):
            assert n == int(torch.tensor(shp).prod()) and off % 4 == 0
        assert total >= count


def test_bad_arguments_are_reported_not_crashed():
    L = _lib.load()
    tot = ctypes.c_longlong()
    assert L.pcg_mnist_layout(7, 64, 6, -1, None, None, ctypes.byref(tot)) == -1
    assert b"net must be" in L.pcg_last_error()


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pcg_b200.mnist.models.generator import ResidualGenerator
    G = ResidualGenerator(base_ch=16, n_resblocks=1)
    x = torch.zeros(2, 1, 28, 28)
    with pytest.raises(RuntimeError):
        G(x, torch.zeros(2, dtype=torch.long), torch.ones_like(x))


@pytest.mark.reference
@pytest.mark.skipif(not have_reference(), reason="no reference")
def test_mirror_modules_match_reference_state_dict_and_init():
    from pcg_b200.mnist.models.generator import ResidualGenerator
    from pcg_b200.mnist.models.discriminator import Discriminator
    from pcg_b200.mnist.models.classifier import CNNClassifier
    with experiment("conditional_counteRGAN/mnist") as imp:
        rg, rd, rc = imp("models.generator"), imp("models.discriminator"), imp("models.classifier")
        pairs = ((ResidualGenerator, rg.ResidualGenerator), (Discriminator, rd.Discriminator),
                 (CNNClassifier, rc.CNNClassifier))
        for mine_cls, ref_cls in pairs:
            torch.manual_seed(11)
            mine = mine_cls()
            torch.manual_seed(11)
            ref = ref_cls()
            sm, sr = mine.state_dict(), ref.state_dict()
            assert list(sm.keys()) == list(sr.keys())
            for k in sr:
                assert sm[k].shape == sr[k].shape and torch.equal(sm[k], sr[k]), k
        # the shipped generator checkpoint loads
        ck = torch.load("/root/reference/conditional_counteRGAN/mnist/results/generator.pt", map_location="cpu")
        ResidualGenerator().load_state_dict(ck)


def test_mlp_gan_step_rejects_shapes_outside_its_envelope():
    """pcg_mlp_gan_step validates before it touches the device: error code + message, no crash (and no fallback)."""
    L = _lib.load()
    f = ctypes.c_float(1e-3)
    args = [None] * 6 + [None] * 10
    assert L.pcg_mlp_gan_step(64, 32, 2, 64, *args, f, None, None) != 0
    assert b"hidden width must be 128" in L.pcg_last_error()
    assert L.pcg_mlp_gan_step(64, 32, 3, 128, *args, f, None, None) != 0
    assert b"label_dim <= 2" in L.pcg_last_error()
    assert L.pcg_mlp_gan_step(64, 40, 2, 128, *args, f, None, None) != 0
    assert b"z_dim + label_dim <= 36" in L.pcg_last_error()


def test_flat_parameter_layout_is_the_one_the_fused_mlp_gan_kernel_indexes():
    """csrc/mlp_gan.cu addresses the caller's flat buffers as [0.weight | 0.bias | 2.weight | 2.bias] with every slice
    padded to a multiple of 4 floats: ob1 = r4(H * in), ow2 = ob1 + H, ob2 = ow2 + out * H, total = ob2 + 4."""
    from pcg_b200.ops import FlatParams
    r4 = lambda n: (n + 3) // 4 * 4  # noqa: E731
    H = 128
    for in_dim, out_dim in ((34, 2), (32, 2), (4, 1), (2, 1)):
        fp = FlatParams([("net.0.weight", (H, in_dim)), ("net.0.bias", (H,)), ("net.2.weight", (out_dim, H)),
                         ("net.2.bias", (out_dim,))], "cpu")
        ob1 = r4(H * in_dim)
        ow2 = ob1 + H
        ob2 = ow2 + out_dim * H
        assert [fp.offsets[k] for k in fp.names] == [0, ob1, ow2, ob2]
        assert fp.size == ob2 + 4
