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


def lift(rel_path, names=(), loop_var=None, extra=None):
    """AST-lifts definitions out of a reference *script* (scripts cannot be imported: they train at import
    time).  Returns (namespace, loop_code): the namespace holds the class / function definitions named in
    ``names`` executed unmodified; ``loop_code`` is the compiled top-level ``for <loop_var> in ...`` statement
    (the script's training loop), or None.  Nothing is copied into the repository."""
    import ast
    import numpy as np
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    src = open(os.path.join(REF_ROOT, rel_path)).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "device": torch.device("cpu")}
    ns.update(extra or {})
    loop = None
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names:
            exec(compile(ast.Module([node], []), rel_path, "exec"), ns)
        if loop_var and isinstance(node, ast.For) and getattr(node.target, "id", None) == loop_var and loop is None:
            loop = compile(ast.Module([node], []), rel_path, "exec")
    return ns, loop
