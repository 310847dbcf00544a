"""Mirror of ``conditional_counteRGAN/mnist/models/discriminator.py`` (Discriminator, :5-38)."""
import torch
import torch.nn as nn

from ._native import NativeNet


class Discriminator(NativeNet):
    _net = 1

    def __init__(self, img_shape=(1, 28, 28), num_classes=10):
        super().__init__()
        C, H, W = img_shape
        if (C, H, W) != (1, 28, 28) or num_classes != 10:
            raise ValueError("the native discriminator is built for img_shape=(1,28,28), num_classes=10")
        self.cond_embed = nn.Embedding(num_classes, H * W)
        self.img_channel = 2
        self.d_hidden = 64
        h = self.d_hidden
        layers = []
        for cin, cout in ((self.img_channel, h), (h, 2 * h), (2 * h, 4 * h), (4 * h, 4 * h)):
            layers += [nn.Conv2d(cin, cout, 3, 2, 1, bias=False), nn.LeakyReLU(0.2, inplace=True)]
        layers.append(nn.AdaptiveAvgPool2d(1))
        self.main = nn.Sequential(*layers)          # conv weights at main.0 / .2 / .4 / .6
        self.flatten = nn.Flatten()
        self.adv_head = nn.Linear(4 * h, 1)

    def forward(self, x, cond_idx):
        """Logits ``[B, 1]`` (discriminator.py:33-38); no autograd graph is recorded."""
        plan = self._plan_for(x.shape[0])
        with torch.no_grad():
            return plan.d_forward(self._img(x), cond_idx.to(torch.int64).contiguous())
