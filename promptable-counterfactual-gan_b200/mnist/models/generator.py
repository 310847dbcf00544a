"""Mirror of ``conditional_counteRGAN/mnist/models/generator.py`` (ResidualGenerator, :25-86).

Same constructor, ``forward`` signature / return values, initialisation and ``state_dict`` keys; the
arithmetic runs in libpcg (sm_100a kernels), not in torch.
"""
import torch
import torch.nn as nn

from ._native import NativeNet


class _ResBlock(nn.Module):
    """Parameter container for conv1-bn1-act-conv2-bn2 with ``x + 0.1*out`` (generator.py:5-22)."""

    def __init__(self, channels):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(channels)
        self.conv2 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn2 = nn.BatchNorm2d(channels)


class ResidualGenerator(NativeNet):
    _net = 0

    def __init__(self, img_shape=(1, 28, 28), num_classes=10, base_ch=64, n_resblocks=6, residual_scaling=0.1):
        super().__init__()
        C, H, W = img_shape
        if (C, H, W) != (1, 28, 28) or num_classes != 10:
            raise ValueError("the native generator is built for img_shape=(1,28,28), num_classes=10")
        if base_ch % 4 or 256 % (base_ch // 4) or n_resblocks < 1:
            raise ValueError("base_ch must be 4*2^k (<=1024) and n_resblocks >= 1")
        self.embed = nn.Embedding(num_classes, H * W)
        self.conv_in = nn.Conv2d(C + 2, base_ch, kernel_size=3, padding=1)
        self.resblocks = nn.Sequential(*[_ResBlock(base_ch) for _ in range(n_resblocks)])
        self.conv_mid = nn.Conv2d(base_ch, base_ch, kernel_size=3, padding=1)
        self.conv_out = nn.Conv2d(base_ch, 1, kernel_size=3, padding=1)
        self.residual_scaling = residual_scaling
        self._init_weights()

    def _init_weights(self):
        # generator.py:58-69 — Kaiming(a=0.2) convs, zero biases, BN (1,0), embedding N(0, 0.01)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, a=0.2)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, mean=0.0, std=0.01)

    def forward(self, x, target, mask=None):
        """Returns ``(raw_residual, masked_residual)`` like generator.py:71-86.  Train mode uses batch
        statistics and updates the BatchNorm running buffers; eval mode uses the running statistics.
        No autograd graph is recorded: gradients are produced by ``trainer.train_countergan``."""
        if mask is None:
            # the reference concatenates ``mask`` into the input (generator.py:74) and fails on None
            raise TypeError("mask is required (it is an input channel of the generator)")
        if abs(self.residual_scaling - 0.1) > 1e-12:
            raise ValueError("native plan is built with residual_scaling=0.1")
        plan = self._plan_for(x.shape[0])
        with torch.no_grad():
            return plan.g_forward(self._img(x), target.to(torch.int64).contiguous(), self._img(mask), self.training)
