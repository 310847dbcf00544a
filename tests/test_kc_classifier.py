"""KC house-sales classifier pre-training (house_sales_kc_usa/trainer.py:18-180,
models/nn_classifier.py:4-32).

CPU : the oracle (oracle/kc_classifier.py) against the reference's own NNClassifier driven by the reference's training
      lines (trainer.py:58-60,85-96: CrossEntropyLoss(weight), AdamW, train-mode BatchNorm; dropout probabilities set to 0
      through the modules' own attribute so both sides are deterministic);
GPU : the native plan against the oracle with injected dropout masks (loss, accuracy counter, every gradient, the AdamW
      update, BatchNorm running buffers, the eval pass), and the drop-in train_classifier end to end.
"""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import kc_classifier as OK
from tests._refload import experiment


def l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def _data(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 17, generator=g)
    y = (x[:, 0] * 2 + x[:, 3] > 1.4).long() + 2 * (x[:, 5] > 0.5).long()
    return x, y


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    P = OrderedDict()
    for k, s in OK.shapes().items():
        if k.endswith(".weight") and len(s) == 2:
            P[k] = torch.randn(*s, generator=g) * (1.0 / s[1]) ** 0.5
        elif k.split(".")[0] + "." + k.split(".")[1] in OK.BN and k.endswith("weight"):
            P[k] = 1.0 + 0.1 * torch.randn(*s, generator=g)
        else:
            P[k] = 0.05 * torch.randn(*s, generator=g)
    return P


@pytest.mark.reference
def test_oracle_matches_reference_classifier_training_lines():
    x, y = _data(96, 1)
    cw = torch.tensor([0.7, 1.1, 1.6, 0.9])
    with experiment("conditional_counteRGAN/house_sales_kc_usa") as imp:
        NN = imp("models.nn_classifier").NNClassifier
        torch.manual_seed(2)
        ref = NN(17, output_dim=4)
        for m in ref.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        PC = OrderedDict((k, v.detach().clone()) for k, v in ref.named_parameters())
        assert list(PC.keys()) == list(OK.shapes().keys())
        crit = torch.nn.CrossEntropyLoss(weight=cw)                                     # trainer.py:58
        opt = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=1e-4)           # trainer.py:60
        ref.train()
        S = OK.make_state(PC)
        ones = [torch.ones(32, c) for c in OK.DIMS[:3]]
        for i in range(3):                                                              # trainer.py:85-96
            xb, yb = x[i * 32:(i + 1) * 32], y[i * 32:(i + 1) * 32]
            before = {k: v.detach().clone() for k, v in ref.named_parameters()}
            opt.zero_grad()
            loss = crit(ref(xb), yb)
            loss.backward()
            gref = {k: v.grad.clone() for k, v in ref.named_parameters()}
            opt.step()
            lo, _, G = OK.train_step(S, xb, yb, ones, cw)
            assert abs(lo - loss.item()) < (1e-6 if i == 0 else 2e-3) * max(1.0, abs(loss.item()))
            for k, v in ref.named_parameters():
                if i == 0:
                    assert l2(G[k], gref[k]) < 1e-5, k
                # Adam turns a gradient that is analytically zero (the bias of a unit whose LeakyReLU has one sign over the
                # whole batch, behind the mean-free BatchNorm backward) into a +-lr step decided by rounding noise: compare
                # the update with the robust mean of test_mnist_step_gpu, in units of lr
                d_ref, d_or = v.detach() - before[k], S["C"][k].detach() - before[k]
                live = gref[k].abs() > 1e-5 * gref[k].abs().max()              # elements with a real gradient
                assert live.float().mean() > 0.5, k
                assert ((d_ref - d_or).abs()[live].mean() / 1e-3).item() < (0.02 if i == 0 else 0.2), (i, k)
            # keep the two sides on the same trajectory for the next iteration
            with torch.no_grad():
                for k, v in ref.named_parameters():
                    S["C"][k].copy_(v)
    for k, v in ref.named_buffers():
        assert torch.allclose(S["B"][k].float(), v.float(), atol=1e-4, rtol=1e-3), k
    ref.eval()
    with torch.no_grad():
        assert torch.allclose(OK.forward({k: v.detach() for k, v in S["C"].items()}, S["B"], x, training=False), ref(x), atol=1e-5)


@pytest.mark.gpu
def test_native_kc_classifier_step_matches_oracle():
    import pcg_b200  # noqa: F401
    from pcg_b200.tabular.kc_classifier import KcClassifierPlan
    B = 128
    PC = _params(7)
    cw = torch.tensor([0.7, 1.1, 1.6, 0.9])
    S = OK.make_state(PC)
    plan = KcClassifierPlan(B, "cuda", lr=1e-3, wd=1e-4, class_weights=cw)
    plan.C.load(PC)
    plan.refresh()
    for step in range(3):
        x, y = _data(B, 20 + step)
        masks = OK.synth_masks(B, 30 + step)
        before = {k: v.detach().clone() for k, v in S["C"].items()}
        loss, correct, G = OK.train_step(S, x, y, masks, cw)
        sc = plan.step(x.cuda(), y.cuda(), masks=[m.cuda() for m in masks])
        torch.cuda.synchronize()
        tol = 1e-4 if step == 0 else 5e-3
        assert abs(sc[0].item() - loss) <= tol * abs(loss), (step, sc[0].item(), loss)
        if step == 0:
            assert int(sc[1].item()) == correct
            for k in G:
                assert l2(plan.C.g(k), G[k]) < 5e-4, (k, l2(plan.C.g(k), G[k]))
            for k in before:
                d_nat = plan.C.p(k).cpu() - before[k]
                d_or = S["C"][k].detach() - before[k]
                live = G[k].abs() > 1e-5 * G[k].abs().max()         # analytically-zero gradients: Adam steps +-lr on noise
                assert ((d_nat - d_or).abs()[live].mean() / 1e-3).item() < 0.02, k
    for j in range(4):
        assert l2(plan.rm[j], S["B"][OK.BN[j] + ".running_mean"]) < 1e-3
        assert l2(plan.rv[j], S["B"][OK.BN[j] + ".running_var"]) < 1e-3
    x, y = _data(B, 99)
    sc = plan.evaluate(x.cuda(), y.cuda())
    P = {k: v.detach() for k, v in S["C"].items()}
    logits = OK.forward(P, S["B"], x, training=False)
    want = torch.nn.functional.cross_entropy(logits, y, weight=cw).item()
    assert abs(sc[2].item() - want) <= 5e-3 * abs(want) and abs(int(sc[3].item()) - int((logits.argmax(1) == y).sum())) <= 1


@pytest.mark.gpu
def test_kc_train_classifier_drop_in(tmp_path):
    import pcg_b200  # noqa: F401
    from pcg_b200.tabular.kc_classifier import train_classifier
    x, y = _data(1100, 5)
    cfg = {"cuda": "cuda", "seed": 3, "input_dim": 17, "out_dir": str(tmp_path), "clf_batch_size": 128, "clf_epochs": 12,
           "clf_early_stopping": 6, "clf_lr": 2e-3, "clf_model_path": str(tmp_path / "clf_model.pt")}
    model = train_classifier(x[:900].numpy(), x[900:].numpy(), y[:900].numpy(), y[900:].numpy(), None, cfg)
    assert cfg["num_classes"] == 4
    sd = torch.load(cfg["clf_model_path"], map_location="cpu")
    assert list(sd.keys()) == list(model.state_dict().keys()) and "net.2.running_var" in sd
    # the saved weights classify the held-out rows far above chance, evaluated with plain torch on the CPU
    ref = torch.nn.Sequential()
    from pcg_b200.tabular.kc import NNClassifier
    m2 = NNClassifier(17, 4)
    m2.load_state_dict(sd)
    m2.eval()
    with torch.no_grad():
        acc = (m2.net(x[900:]).argmax(1) == y[900:]).float().mean().item()
    assert acc > 0.7, acc
