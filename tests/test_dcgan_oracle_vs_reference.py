"""Pins oracle/dcgan.py against the reference script's own Generator / Discriminator / weights_init classes
(AST-lifted, build container only) and its loop body executed statement by statement as written."""
from collections import OrderedDict

import pytest
import torch

from oracle import dcgan as O
from tests._refload import have_reference, lift

pytestmark = pytest.mark.reference


@pytest.mark.skipif(not have_reference(), reason="no reference")
def test_modules_and_two_iterations_match():
    import torch.nn as nn
    import torch.optim as optim
    cfg = {'batch_size': 4, 'image_channel': 1, 'z_dim': 100, 'g_hidden': 64, 'd_hidden': 64, 'x_dim': 64,
           'real_label': 1., 'fake_label': 0., 'lr': 2e-4}
    ns, _ = lift("dconv_gan/mnist/mnist_dcgan.py", ("Generator", "Discriminator", "weights_init"), extra={"config": cfg})
    torch.manual_seed(3)
    netG, netD = ns["Generator"](), ns["Discriminator"]()
    netG.apply(ns["weights_init"])
    netD.apply(ns["weights_init"])
    assert list(k for k, v in netG.state_dict().items() if v.dim() > 0 and "running" not in k) == list(O.g_shapes().keys())
    assert list(k for k, v in netD.state_dict().items() if v.dim() > 0 and "running" not in k) == list(O.d_shapes().keys())
    S = O.make_state(OrderedDict(netG.named_parameters()), OrderedDict(netG.named_buffers()),
                     OrderedDict(netD.named_parameters()), OrderedDict(netD.named_buffers()))
    criterion = nn.BCELoss()
    optD = optim.Adam(netD.parameters(), lr=cfg['lr'], betas=(0.5, 0.999))
    optG = optim.Adam(netG.parameters(), lr=cfg['lr'], betas=(0.5, 0.999))
    for it in range(2):
        real, noise = O.synth_batch(4, 11 + it)
        # ---- the reference loop body, mnist_dcgan.py:147-175, with `noise` injected at :156
        netD.zero_grad()
        label = torch.full((4,), cfg['real_label'], dtype=torch.float)
        output = netD(real)
        errD_real = criterion(output, label)
        errD_real.backward()
        fake = netG(noise)
        label.fill_(cfg['fake_label'])
        output = netD(fake.detach())
        errD_fake = criterion(output, label)
        errD_fake.backward()
        errD = errD_real + errD_fake
        optD.step()
        netG.zero_grad()
        label.fill_(cfg['real_label'])
        output = netD(fake)
        errG = criterion(output, label)
        errG.backward()
        optG.step()
        sc, _ = O.dcgan_step(S, real, noise)
        assert abs(sc["errD"] - errD.item()) < 2e-5 * abs(errD.item()) + 1e-6
        assert abs(sc["errG"] - errG.item()) < 2e-5 * abs(errG.item()) + 1e-6
    def close(v, mine, k):
        v, mine = v.float(), mine.detach().float()
        if "running" in k or "num_batches" in k:
            assert torch.allclose(v, mine, atol=1e-5, rtol=1e-3), k
        else:       # Adam: an element whose gradient is rounding noise may move +-lr on either side
            d = (v - mine).abs()
            assert d.max() <= 2.02 * cfg['lr'] * 2 and d.mean() <= 0.02 * cfg['lr'], (k, d.max(), d.mean())
    for k, v in netG.state_dict().items():
        close(v, S["G"][k] if k in S["G"] else S["GB"][k], k)
    for k, v in netD.state_dict().items():
        close(v, S["D"][k] if k in S["D"] else S["DB"][k], k)
