"""Classifier pre-training (conditional_counteRGAN/mnist/trainer.py:8-39, models/classifier.py:4-28).

CPU  : the oracle (oracle/mnist_classifier.py) against the UNMODIFIED reference train_classifier run here (dropout
       probabilities of the reference module set to 0 through its own attributes, so both sides are deterministic);
GPU  : the native plan (pcg_b200/mnist/classifier_trainer.py) against the oracle with injected dropout masks - loss, every
       gradient, the Adam updates, the eval forward - the statistics of the dropout-mask kernel, and the drop-in
       train_classifier end to end.
"""
import types
from collections import OrderedDict

import pytest
import torch

from oracle import mnist_classifier as OC
from oracle import mnist_countergan as O
from tests._refload import experiment


def l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def _batches(n, B, seed=0):
    out = []
    for i in range(n):
        x, y, _, _ = O.synth_batch(B, seed + i)
        out.append((x, y))
    return out


@pytest.mark.reference
def test_oracle_matches_reference_train_classifier(tmp_path):
    batches = _batches(4, 8, 40)
    with experiment("conditional_counteRGAN/mnist") as imp:
        trainer, Cm = imp("trainer"), imp("models.classifier")
        torch.manual_seed(3)
        ref = Cm.CNNClassifier()
        for m in ref.modules():
            if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)):
                m.p = 0.0                       # the module's own attribute: no random draw on either side
        PC = OrderedDict((k, v.detach().clone()) for k, v in ref.named_parameters())
        cfg = types.SimpleNamespace(cls_lr=1e-3, num_epochs_clf=2, classifier_path=str(tmp_path / "c.pt"))
        trainer.train_classifier(ref, batches[:3], batches[3:], cfg, "cpu")
    S = OC.make_state(PC)
    ones2, ones1 = torch.ones(8, 128), torch.ones(8, 256)
    for _ in range(2):
        for x, y in batches[:3]:
            OC.train_step(S, x, y, ones2, ones1, lr=1e-3)
    for k, v in ref.named_parameters():
        assert torch.allclose(S["C"][k].detach(), v.detach(), atol=2e-6, rtol=1e-4), k
    # the masks enter exactly where and how nn.Dropout2d / nn.Dropout do
    m2, m1 = OC.synth_masks(8, 5)
    x = batches[0][0]
    P = OrderedDict((k, v.detach()) for k, v in ref.named_parameters())
    z = ref.conv[:6](x) * m2.view(8, 128, 1, 1)
    want = ref.fc[4](torch.relu(ref.fc[1](z.flatten(1))) * m1)
    assert torch.allclose(OC.forward_train(P, x, m2, m1), want, atol=1e-6)


@pytest.mark.gpu
def test_native_classifier_step_matches_oracle():
    import pcg_b200  # noqa: F401
    from pcg_b200.mnist.classifier_trainer import ClassifierPlan
    B = 16
    PC = O.synth_params(O.c_param_shapes(), 3, "C")
    S = OC.make_state(PC)
    for graph in (False, True):
        S = OC.make_state(PC)
        plan = ClassifierPlan(B, "cuda", lr=1e-3, use_graph=graph)
        plan.C.load(PC)
        plan.refresh()
        for step in range(3):
            x, y, _, _ = O.synth_batch(B, 60 + step, mnist_like=(step == 1))
            m2, m1 = OC.synth_masks(B, 90 + step)
            before = {k: v.detach().clone() for k, v in S["C"].items()}
            loss, G = OC.train_step(S, x, y, m2, m1, lr=1e-3)
            got = plan.step(x.cuda(), y.cuda(), masks=(m2.cuda(), m1.cuda()))
            torch.cuda.synchronize()
            tol = 1e-4 if step == 0 else 5e-3
            assert abs(got.item() - loss) <= tol * abs(loss), (graph, step, got.item(), loss)
            if step == 0:
                for k in G:
                    assert l2(plan.C.g(k), G[k]) < 2e-4, (k, l2(plan.C.g(k), G[k]))
                for k in before:                # Adam update in units of lr (robust mean, see test_mnist_step_gpu)
                    d_nat = plan.C.p(k).cpu() - before[k]
                    d_or = S["C"][k].detach() - before[k]
                    assert ((d_nat - d_or).abs().mean() / 1e-3).item() < 0.02, k
        x, _, _, _ = O.synth_batch(B, 77)
        assert l2(plan.logits(x.cuda()), O.c_forward({k: v.detach() for k, v in S["C"].items()}, x)) < 5e-3


@pytest.mark.gpu
def test_dropout_mask_kernel_statistics():
    import pcg_b200  # noqa: F401
    from pcg_b200 import ops as K
    rng = torch.zeros(3, dtype=torch.int64, device="cuda")
    m = torch.empty(4096, 256, device="cuda")
    K.dropout_mask(m, 0.5, rng_state=rng)
    assert set(m.unique().tolist()) == {0.0, 2.0} and abs((m > 0).float().mean().item() - 0.5) < 5e-3
    assert abs((m[:, :128] * m[:, 128:]).mean().item() - 1.0) < 0.02           # independent draws
    m2 = torch.empty(2048, 7, 7, 128, device="cuda")
    K.dropout_mask(m2, 0.25, channelwise=True, rng_state=rng)
    keep = 1.0 / 0.75
    assert torch.equal(m2, m2[:, :1, :1, :].expand_as(m2))                     # whole feature maps, per sample
    assert abs((m2[:, 0, 0] > 0).float().mean().item() - 0.75) < 6e-3 and abs(m2.max().item() - keep) < 1e-6
    again = torch.empty_like(m)
    K.dropout_mask(again, 0.5, rng_state=rng)
    assert not torch.equal(again, m) and int(rng[0]) == 3                      # every launch advances the stream


@pytest.mark.gpu
def test_train_classifier_drop_in(tmp_path):
    """Same signature and side effects as trainer.py:8-39; learns a separable synthetic task (class = brightest of ten
    fixed patches), ragged last batch included."""
    import pcg_b200  # noqa: F401
    from pcg_b200.mnist.classifier_trainer import train_classifier
    from pcg_b200.mnist.models.classifier import CNNClassifier
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1)

    def make(n):
        y = torch.randint(0, 10, (n,), generator=g)
        x = torch.randn(n, 1, 28, 28, generator=g) * 0.3 - 0.5
        for i, c in enumerate(y.tolist()):
            r, col = divmod(c, 5)
            x[i, 0, 4 + 12 * r: 12 + 12 * r, 1 + 5 * col: 6 + 5 * col] += 1.5
        return x, y
    xt, yt = make(600)
    xv, yv = make(200)
    train = [(xt[i:i + 64], yt[i:i + 64]) for i in range(0, 600, 64)]          # last batch: 24 samples
    valid = [(xv[i:i + 64], yv[i:i + 64]) for i in range(0, 200, 64)]
    C = CNNClassifier()
    cfg = types.SimpleNamespace(cls_lr=1e-3, num_epochs_clf=3, classifier_path=str(tmp_path / "best_classifier.pt"))
    acc = train_classifier(C, train, valid, cfg, "cuda")
    assert acc > 0.9, acc
    sd = torch.load(cfg.classifier_path, map_location="cpu")
    assert list(sd.keys()) == list(C.state_dict().keys())
    C.eval()
    with torch.no_grad():
        pred = torch.cat([C(x.cuda()).argmax(1).cpu() for x, _ in valid])
    assert (pred == yv).float().mean().item() > 0.85                           # the module sees the trained weights
