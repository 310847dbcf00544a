"""MLP GANs on 2-D points: conditional_gan/moons/make_moons_cgan.py and simple_gan/moons/make_moons_gan.py."""
