"""CPU oracle for the GAN training step of flash4242/Promptable-Counterfactual-GAN.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as the
CPU arm being timed), never as the thing shipped.

The reference is pure Python on top of PyTorch, so the arithmetic "lives" in a
third-party dependency (torch; pinned by the reference's Dockerfile:1 at
2.4.1+cu118, present here as 2.11.0+cu128).  The oracle therefore restates

  * the reference's own composition of that arithmetic (layer order, losses,
    the D-step / G-step ordering, what is detached, which D weights the G step
    sees) as plain functions over dictionaries of tensors keyed by the
    reference's ``state_dict`` names, and
  * the published algorithms of the torch ops whose semantics matter for parity
    (train-mode BatchNorm incl. running-stat update, BCE-with-logits,
    cross-entropy, Adam) explicitly, in fp32 (or fp64) tensor arithmetic,

and uses ``torch.nn.functional.conv2d`` / autograd on CPU for the convolutions and
their gradients (the contraction itself has no semantics to restate).

Pinning: ``tests/test_oracle_vs_reference.py`` runs the reference's *unmodified*
``train_countergan`` / modules imported from ``/root/reference`` (only where that
tree exists, i.e. in the build container) against this oracle, and
``tests/golden/`` holds vectors generated from the reference by
``tests/golden/make_golden.py``.  The reference itself ships no tests or golden
vectors for this path.
"""
