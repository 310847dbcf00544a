"""Conditional WGAN-GP (conditional_gan/mnist/mnist_wgan_conditional.py) on libpcg operators."""
from .plan import Critic, Generator, Hyperparameter, WganGpPlan, train_wgan_gp  # noqa: F401
