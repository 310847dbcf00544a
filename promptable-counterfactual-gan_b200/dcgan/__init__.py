"""DCGAN on 1x64x64 images: dconv_gan/mnist/mnist_dcgan.py."""
from .plan import DcganPlan, Generator, Discriminator, weights_init, train_dcgan, config  # noqa: F401
