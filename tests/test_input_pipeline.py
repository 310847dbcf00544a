"""On-device input pipeline: the normalisation formula is pinned to torchvision's own transform
on CPU; the native kernel must reproduce it bit for bit, through the C ABI, for shuffled and ragged batches."""
import numpy as np
import pytest
import torch


def formula(u8):
    """transforms.ToTensor() (float / 255) followed by transforms.Normalize((0.5,), (0.5,)), data_utils.py:9-12."""
    return (u8.float().div(255) - 0.5) / 0.5


def test_formula_is_torchvisions_transform():
    tv = pytest.importorskip("torchvision")
    PIL = pytest.importorskip("PIL.Image")
    from torchvision import transforms
    t = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5,), (0.5,))])
    rng = np.random.RandomState(0)
    imgs = rng.randint(0, 256, (8, 28, 28)).astype(np.uint8)
    imgs[0] = np.arange(784).reshape(28, 28) % 256          # every byte value
    for im in imgs:
        got = t(PIL.fromarray(im, mode="L"))
        assert torch.equal(got, formula(torch.from_numpy(im)).unsqueeze(0))


def test_stratified_split_matches_reference_call():
    from sklearn.model_selection import train_test_split
    y = np.random.RandomState(1).randint(0, 10, 1000)
    a = train_test_split(list(range(1000)), test_size=0.1, stratify=y, random_state=3)
    import importlib
    du = importlib.import_module("pcg_b200.mnist.data_utils") if _has_pkg() else None
    if du is None:
        pytest.skip("package import needs the built library")
    # same call, same seed -> same split (the loaders only move the index sets to the device)
    b = train_test_split(list(range(1000)), test_size=0.1, stratify=y, random_state=3)
    assert a[0] == b[0] and a[1] == b[1] and len(a[1]) == 100


def _has_pkg():
    try:
        import pcg_b200  # noqa: F401
        return True
    except Exception:
        return False


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 7, 512])
def test_u8_batch_is_bit_exact(B):
    import pcg_b200  # noqa: F401
    from pcg_b200.mnist.data_utils import u8_batch
    g = torch.Generator().manual_seed(5)
    imgs = torch.randint(0, 256, (1000, 28, 28), generator=g, dtype=torch.uint8)
    imgs[0] = (torch.arange(784) % 256).to(torch.uint8).view(28, 28)
    labels = torch.randint(0, 10, (1000,), generator=g)
    idx = torch.randperm(1000, generator=g)[:B]
    idx[0] = 0
    x, y = u8_batch(imgs.cuda(), labels.cuda(), idx.cuda())
    assert x.shape == (B, 1, 28, 28) and torch.equal(x.cpu(), formula(imgs[idx]).unsqueeze(1))
    assert torch.equal(y.cpu(), labels[idx])
    x2, _ = u8_batch(imgs[:B].contiguous().cuda(), None, None)          # identity gather, no labels
    assert torch.equal(x2.cpu(), formula(imgs[:B]).unsqueeze(1))


@pytest.mark.gpu
def test_device_loader_epoch_and_trainer_consumes_it():
    import types
    import pcg_b200  # noqa: F401
    from pcg_b200.mnist import data_utils as DU
    g = torch.Generator().manual_seed(6)
    tr_u8 = torch.randint(0, 256, (300, 28, 28), generator=g, dtype=torch.uint8)
    tr_y = torch.arange(300) % 10
    te_u8 = torch.randint(0, 256, (50, 28, 28), generator=g, dtype=torch.uint8)
    te_y = torch.arange(50) % 10
    train, valid, test = DU.loaders_from_tensors(tr_u8, tr_y, te_u8, te_y, batch_size=64, seed=0)
    assert len(train.dataset) == 270 and len(valid.dataset) == 30 and len(test.dataset) == 50 and len(train) == 5
    seen, n = [], 0
    for x, y in train:                                       # one shuffled epoch covers the training subset exactly once
        assert x.is_cuda and x.dtype == torch.float32 and x.min() >= -1 and x.max() <= 1
        n += x.shape[0]
        seen.append(y)
    assert n == 270
    counts = torch.bincount(torch.cat(seen).cpu(), minlength=10)
    assert counts.tolist() == [27] * 10                      # stratified 90 % of 30 per class
    xs = torch.cat([x for x, _ in test]).cpu()
    assert torch.equal(xs, formula(te_u8).unsqueeze(1))      # unshuffled loader = dataset order
    # the trainer consumes the loader unchanged
    from pcg_b200.mnist.models.classifier import CNNClassifier
    from pcg_b200.mnist.models.discriminator import Discriminator
    from pcg_b200.mnist.models.generator import ResidualGenerator
    from pcg_b200.mnist import trainer as T
    torch.manual_seed(0)
    G, D, C = ResidualGenerator().cuda(), Discriminator().cuda(), CNNClassifier().cuda().eval()
    cfg = types.SimpleNamespace(g_lr=5e-5, d_lr=1e-5, num_classes=10, patch_size=7, num_modifiable_patches=10,
                                lambda_adv=1.0, lambda_cls=1.0, lambda_reg=2.5, lambda_mask=2.0)
    tr = T.CounterGanTrainer(G, D, C, cfg, "cuda")
    for x, y in DU.DeviceLoader(tr_u8.cuda(), tr_y.cuda(), None, 32, shuffle=True):
        tgt = torch.randint(0, 10, (x.shape[0],), device="cuda")
        p = tr.step(x.contiguous(), y, tgt, T.build_mask(x, 7, "cuda", 10).contiguous())
        break
    s = p.scalars_dict()
    assert all(np.isfinite(v) for v in s.values())
