"""Loads the UNMODIFIED reference experiment modules from /root/reference (build container only).

Each reference experiment uses bare names (``config``, ``models``, ``trainer``) and must have its own
directory first on ``sys.path`` (conditional_counteRGAN/mnist/main.py:4-10); ``trainer.py`` imports
matplotlib at module top (trainer.py:5), which is not installed here, so it is stubbed.
Nothing is copied: the modules are imported in place.
"""
import contextlib
import importlib
import os
import sys
from unittest import mock

REF_ROOT = "/root/reference"
_BARE = ("config", "models", "trainer", "data_utils", "eval_utils")


def have_reference():
    return os.path.isdir(os.path.join(REF_ROOT, "conditional_counteRGAN"))


def _purge():
    for k in list(sys.modules):
        if k in _BARE or k.startswith("models."):
            del sys.modules[k]


@contextlib.contextmanager
def experiment(rel_dir):
    """Context: the reference experiment at ``rel_dir`` importable by its bare module names."""
    d = os.path.join(REF_ROOT, rel_dir)
    _purge()
    for m in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules.setdefault(m, mock.MagicMock())
    sys.path.insert(0, d)
    cwd = os.getcwd()
    try:
        yield lambda name: importlib.import_module(name)
    finally:
        os.chdir(cwd)
        sys.path.remove(d)
        _purge()
