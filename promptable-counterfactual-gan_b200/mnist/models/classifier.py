"""Mirror of ``conditional_counteRGAN/mnist/models/classifier.py`` (CNNClassifier, :4-28)."""
import torch
import torch.nn as nn

from ._native import NativeNet


class CNNClassifier(NativeNet):
    _net = 2

    def __init__(self, num_classes=10):
        super().__init__()
        if num_classes != 10:
            raise ValueError("the native classifier is built for num_classes=10")
        self.conv = nn.Sequential(
            nn.Conv2d(1, 32, 3, 1, 1), nn.ReLU(),
            nn.Conv2d(32, 64, 3, 2, 1), nn.ReLU(),
            nn.Conv2d(64, 128, 3, 2, 1), nn.ReLU(),
            nn.Dropout2d(0.25))
        self.fc = nn.Sequential(nn.Flatten(), nn.Linear(128 * 7 * 7, 256), nn.ReLU(), nn.Dropout(0.5),
                                nn.Linear(256, num_classes))

    def forward(self, x):
        """Logits ``[B, 10]`` with dropout inactive — the frozen, eval-mode classifier of the hot path
        (main.py:31-33).  Train-mode dropout belongs to classifier pre-training, which is out of scope."""
        if self.training:
            raise NotImplementedError("native CNNClassifier.forward implements the frozen eval-mode classifier; "
                                      "call .eval() (classifier pre-training is outside the hot path)")
        plan = self._plan_for(x.shape[0])
        with torch.no_grad():
            return plan.c_forward(self._img(x))
