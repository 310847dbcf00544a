"""Tabular CounteRGANs: conditional_counteRGAN/moons and conditional_counteRGAN/house_sales_kc_usa."""
