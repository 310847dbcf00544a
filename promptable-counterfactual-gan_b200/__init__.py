"""B200-native GAN training step for Promptable-Counterfactual-GAN (hot path only).

The directory name carries a hyphen (it mirrors the reference repository's name), so import it
through the ``pcg_b200`` shim at the repository root: ``import pcg_b200``.
"""
from . import _lib  # noqa: F401
