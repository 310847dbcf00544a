"""GPU: the drop-in surface — mirror modules' forward and ``train_countergan`` — against the oracle."""
import types
from collections import OrderedDict

import pytest
import torch

from oracle import mnist_countergan as O

pytestmark = pytest.mark.gpu


def _mods(ch=16, nres=2, seed=0):
    import pcg_b200  # noqa: F401
    from pcg_b200.mnist.models.generator import ResidualGenerator
    from pcg_b200.mnist.models.discriminator import Discriminator
    from pcg_b200.mnist.models.classifier import CNNClassifier
    torch.manual_seed(seed)
    G = ResidualGenerator(base_ch=ch, n_resblocks=nres)
    D = Discriminator()
    C = CNNClassifier().eval()
    for p in C.parameters():
        p.requires_grad = False
    # perturb BN affine so they are not (1, 0)
    with torch.no_grad():
        for n, p in G.named_parameters():
            if ".bn" in n:
                p.add_(0.1 * torch.randn_like(p))
    return G, D, C


def _cpu_state(G, D, C):
    return O.make_state(OrderedDict((k, v.detach().cpu()) for k, v in G.named_parameters()),
                        OrderedDict((k, v.detach().cpu().clone()) for k, v in G.named_buffers()),
                        OrderedDict((k, v.detach().cpu()) for k, v in D.named_parameters()),
                        OrderedDict((k, v.detach().cpu()) for k, v in C.named_parameters()))


@pytest.mark.parametrize("precision,tol", [("fp32", 3e-5), ("bf16", 3e-2)])
def test_module_forwards(precision, tol):
    G, D, C = _mods(64 if precision == "bf16" else 16, 2)
    S = _cpu_state(G, D, C)
    G, D, C = G.cuda(), D.cuda(), C.cuda()
    for m in (G, D, C):
        m.precision = precision
    x, y, t, mask = O.synth_batch(8, 3)
    with torch.no_grad():
        raw_o, masked_o = O.g_forward(S["G"], S["GB"], x, t, mask, n_resblocks=2, training=True)
        raw, masked = G(x.cuda(), t.cuda(), mask.cuda())
        rel = lambda a, b: ((a.cpu() - b).abs().max() / b.abs().max()).item()  # noqa: E731
        assert raw.shape == (8, 1, 28, 28) and rel(raw, raw_o) < tol and rel(masked, masked_o) < tol
        # running stats were updated through the module's own buffers
        rm = G.resblocks[1].bn2.running_mean
        assert rel(rm, S["GB"]["resblocks.1.bn2.running_mean"]) < max(tol, 1e-4)
        assert int(G.resblocks[0].bn1.num_batches_tracked) == 1
        # eval mode uses the running statistics
        G.eval()
        raw_e, _ = G(x.cuda(), t.cuda(), mask.cuda())
        raw_eo, _ = O.g_forward(S["G"], S["GB"], x, t, mask, n_resblocks=2, training=False)
        assert rel(raw_e, raw_eo) < tol * 3
        assert rel(D(x.cuda(), y.cuda()), O.d_forward(S["D"], x, y)) < tol * 3
        assert rel(C(x.cuda()), O.c_forward(S["C"], x)) < tol * 3
        # state_dict round trip through the arena views, then a changed weight is seen by the kernels
        sd = {k: v.clone() for k, v in D.state_dict().items()}
        sd["adv_head.bias"] += 1.0
        before = D(x.cuda(), y.cuda())
        D.load_state_dict(sd)
        after = D(x.cuda(), y.cuda())
        assert torch.allclose(after, before + 1.0, atol=1e-2 if precision == "bf16" else 1e-5)
    with pytest.raises(TypeError):
        G(x.cuda(), t.cuda(), None)


@pytest.mark.parametrize("use_graph", [False, True])
def test_train_countergan_matches_oracle(tmp_path, use_graph, monkeypatch):
    from pcg_b200.mnist import trainer as T
    monkeypatch.setenv("PCG_PRECISION", "fp32")
    monkeypatch.setenv("PCG_NO_GRAPH", "0" if use_graph else "1")
    G, D, C = _mods(16, 2, seed=5)
    S = _cpu_state(G, D, C)
    G, D, C = G.cuda(), D.cuda(), C.cuda()
    B, n_steps = 8, 3
    batches = [O.synth_batch(B, 900 + i, mnist_like=(i == 1)) for i in range(n_steps)]
    cfg = types.SimpleNamespace(g_lr=5e-5, d_lr=1e-5, num_epochs_gan=1, num_classes=10, patch_size=7,
                                num_modifiable_patches=10, lambda_adv=1.0, lambda_cls=1.0, lambda_reg=2.5,
                                lambda_mask=2.0, save_dir=str(tmp_path), generator_path=str(tmp_path / "generator.pt"))
    it = {"i": 0}

    def fake_draw(x, cfg_, device):          # the reference's draws (trainer.py:94-95), injected
        t, m = batches[it["i"]][2].clone().to(device), batches[it["i"]][3].clone().to(device)
        it["i"] += 1
        return t, m

    monkeypatch.setattr(T, "draw_target_and_mask", fake_draw)
    out = T.train_countergan(G, D, C, [(b[0], b[1]) for b in batches], cfg, "cuda")
    sums = {"g_loss": 0.0, "d_loss": 0.0, "g_cls": 0.0}
    for (x, y, t, m) in batches:
        sc, _ = O.countergan_step(S, x, y, t, m, n_resblocks=2)
        for k in sums:
            sums[k] += sc[k]
    assert abs(out["g_losses"][0] - sums["g_loss"] / n_steps) < 2e-3 * abs(sums["g_loss"] / n_steps)
    assert abs(out["d_losses"][0] - sums["d_loss"] / n_steps) < 2e-3 * abs(sums["d_loss"] / n_steps)
    saved = torch.load(cfg.generator_path, map_location="cpu")
    assert list(saved.keys()) == list(G.state_dict().keys())
    for k, v in saved.items():
        ref = (S["G"][k] if k in S["G"] else S["GB"][k]).detach().float()
        if O.is_bn_shadowed_bias(k):
            assert (v.float() - ref).abs().max() <= cfg.g_lr * n_steps * 2.02
        elif k in S["G"]:
            # mean absolute deviation of the cumulative update, in units of lr (see test_mnist_step_gpu)
            assert ((v.float() - ref).abs().mean() / cfg.g_lr).item() < 0.2, k
        else:
            assert torch.allclose(v.float(), ref, rtol=2e-3, atol=1e-5), k
    assert int(saved["resblocks.0.bn1.num_batches_tracked"]) == n_steps


def test_train_countergan_ragged_tail_batch_two_epochs(tmp_path, monkeypatch):
    """The reference loader has no drop_last (data_utils.py:27): every epoch ends with a smaller batch, so the trainer
    alternates between two native plans.  Each plan keeps its own packed conv weights; they must be re-packed when the
    other plan updated the parameters.  Two epochs of batches 8, 8, 5 against the oracle."""
    from pcg_b200.mnist import trainer as T
    monkeypatch.setenv("PCG_PRECISION", "fp32")
    monkeypatch.setenv("PCG_NO_GRAPH", "0")
    G, D, C = _mods(16, 2, seed=7)
    S = _cpu_state(G, D, C)
    G, D, C = G.cuda(), D.cuda(), C.cuda()
    sizes = [8, 8, 5]
    epoch = [O.synth_batch(b, 700 + i) for i, b in enumerate(sizes)]
    draws = [O.synth_batch(b, 800 + i) for i, b in enumerate(sizes * 2)]      # targets / masks of the 6 iterations
    # large steps: a stale weight copy (one epoch old) must be visible above the tolerances
    cfg = types.SimpleNamespace(g_lr=2e-3, d_lr=2e-3, num_epochs_gan=2, num_classes=10, patch_size=7,
                                num_modifiable_patches=10, lambda_adv=1.0, lambda_cls=1.0, lambda_reg=2.5,
                                lambda_mask=2.0, save_dir=str(tmp_path), generator_path=str(tmp_path / "generator.pt"))
    it = {"i": 0}

    def fake_draw(x, cfg_, device):
        t, m = draws[it["i"]][2].clone().to(device), draws[it["i"]][3].clone().to(device)
        it["i"] += 1
        return t, m

    monkeypatch.setattr(T, "draw_target_and_mask", fake_draw)
    out = T.train_countergan(G, D, C, [(b[0], b[1]) for b in epoch], cfg, "cuda")
    per_epoch = []
    for e in range(2):
        tot = 0.0
        for j, (x, y, _, _) in enumerate(epoch):
            sc, _ = O.countergan_step(S, x, y, draws[3 * e + j][2], draws[3 * e + j][3], n_resblocks=2,
                                      hp=O.Hyper(g_lr=cfg.g_lr, d_lr=cfg.d_lr))
            tot += sc["g_loss"]
        per_epoch.append(tot / 3)
    for e in range(2):
        assert abs(out["g_losses"][e] - per_epoch[e]) < 5e-3 * abs(per_epoch[e]), (e, out["g_losses"], per_epoch)
    saved = torch.load(cfg.generator_path, map_location="cpu")
    for k, v in saved.items():
        if k in S["G"] and not O.is_bn_shadowed_bias(k):
            assert ((v.float() - S["G"][k].detach().float()).abs().mean() / cfg.g_lr).item() < 0.25, k
    # the module's own forward (forward-only plan cached before training) sees the trained weights
    x, y, t, m = epoch[0]
    G.precision = "fp32"
    G.eval()
    with torch.no_grad():
        raw, _ = G(x.cuda(), t.cuda(), m.cuda())
        raw_o, _ = O.g_forward(S["G"], S["GB"], x, t, m, n_resblocks=2, training=False)
    assert ((raw.cpu() - raw_o).abs().max() / raw_o.abs().max()).item() < 1e-3


def test_build_mask_distribution():
    """Properties of trainer.py:45-72 / :94 the native Philox kernel must reproduce (same distribution, not the same
    stream): exactly num_modifiable_patches * patch^2 ones per sample, every patch equally likely, every PAIR of patches
    equally likely (a uniformly random subset, not just uniform marginals), uniform targets, fresh draws per launch."""
    from pcg_b200.mnist import trainer as T
    import types
    B = 8192
    x = torch.zeros(B, 1, 28, 28, device="cuda")
    torch.manual_seed(0)
    m = T.build_mask(x, 7, "cuda", 10)
    assert m.shape == x.shape and m.dtype == torch.float32
    assert torch.all(m.sum(dim=(1, 2, 3)) == 490)                              # trainer.py:63-65 semantics
    assert torch.equal(m, m.round()) and m.min() == 0 and m.max() == 1
    pm = m[:, 0, ::7, ::7].reshape(B, 16)
    assert torch.equal(torch.nn.functional.interpolate(pm.view(B, 1, 4, 4), size=(28, 28), mode="nearest"), m)
    freq = pm.mean(0)                                                          # each patch chosen w.p. 10/16
    assert torch.all((freq - 10 / 16).abs() < 0.03), freq
    pair = (pm.t() @ pm) / B                                                   # P(i and j) = 10/16 * 9/15 = 0.375
    off = pair[~torch.eye(16, dtype=torch.bool, device="cuda")]
    assert torch.all((off - 0.375).abs() < 0.03), (off.min(), off.max())
    # consecutive launches advance the stream; re-seeding restarts it
    m2 = T.build_mask(x, 7, "cuda", 10)
    assert not torch.equal(m, m2)
    torch.manual_seed(0)
    assert torch.equal(T.build_mask(x, 7, "cuda", 10), m)
    # target draw in the same launch: uniform over the classes, independent of the mask
    cfg = types.SimpleNamespace(patch_size=7, num_modifiable_patches=10, num_classes=10)
    t, m3 = T.draw_target_and_mask(x, cfg, "cuda")
    assert t.dtype == torch.int64 and t.min() >= 0 and t.max() <= 9 and torch.all(m3.sum(dim=(1, 2, 3)) == 490)
    hist = torch.bincount(t, minlength=10).float() / B
    assert torch.all((hist - 0.1).abs() < 0.015), hist
    # None / >= total: independent fair coins per patch (trainer.py:59-61)
    mb = T.build_mask(x, 7, "cuda", None)
    pb = mb[:, 0, ::7, ::7].reshape(B, 16)
    assert torch.all((pb.mean(0) - 0.5).abs() < 0.03) and (pb.sum(1).float().var() - 4.0).abs() < 0.4
    # general geometry: 3 channels, sizes that are not multiples of the patch: nearest up-sampling as F.interpolate
    xg = torch.zeros(33, 3, 30, 33, device="cuda")
    mg = T.build_mask(xg, 7, "cuda", 5)
    ih = [-(-ph * 30 // 4) for ph in range(4)]
    iw = [-(-pw * 33 // 4) for pw in range(4)]
    pg = mg[:, 0][:, ih][:, :, iw]
    assert torch.all(pg.sum(dim=(1, 2)) == 5)
    want = torch.nn.functional.interpolate(pg.unsqueeze(1), size=(30, 33), mode="nearest").repeat(1, 3, 1, 1)
    assert torch.equal(mg, want)


def test_step_auto_draws_inside_the_graph(tmp_path):
    """``step_auto`` (what train_countergan calls): host batches in, target + mask drawn by the first node of the replayed
    graph, fresh at every replay; the losses stay finite and the step counter advances."""
    from pcg_b200.mnist import trainer as T
    G, D, C = _mods(64, 2, seed=3)
    G, D, C = G.cuda(), D.cuda(), C.cuda()
    cfg = types.SimpleNamespace(g_lr=5e-5, d_lr=1e-5, num_classes=10, patch_size=7, num_modifiable_patches=10,
                                lambda_adv=1.0, lambda_cls=1.0, lambda_reg=2.5, lambda_mask=2.0)
    tr = T.CounterGanTrainer(G, D, C, cfg, "cuda", precision="bf16")
    T._warm_plan(tr, 16)
    x, y, _, _ = O.synth_batch(16, 11)
    hx, hy = x.pin_memory(), y.pin_memory()
    masks, targets = [], []
    for _ in range(3):
        p = tr.step_auto(hx, hy)
        torch.cuda.synchronize()
        st = tr.static[16]
        masks.append(st[3].clone())
        targets.append(st[2].clone())
        assert torch.isfinite(p.scalars[:10]).all()
        assert torch.all(st[3].sum(dim=(1, 2, 3)) == 490) and torch.equal(st[0].cpu(), x)
    assert not torch.equal(masks[0], masks[1]) and not torch.equal(masks[1], masks[2])
    assert int(p.adam["g_step"]) == 3
