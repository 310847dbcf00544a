"""CPU: the reference arm of bench.py drives the staged, unmodified reference trainer (baseline/_ref)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_timed_loader_stamps_between_iterations():
    import bench
    ld = bench._TimedLoader([("a", 0), ("b", 1)], warmup=2, steps=3, budget_s=1e9)
    seen = [b for b in ld]
    assert len(seen) == 5 and seen[0] == ("a", 0) and seen[1] == ("b", 1)
    n, dt = ld.timed()
    assert n == 3 and dt >= 0.0 and len(ld.stamps) == 6
    # budget exhausted: never fewer than min_steps timed iterations
    ld = bench._TimedLoader([("a", 0)], warmup=1, steps=50, budget_s=0.0, min_steps=3)
    assert len(list(ld)) == 4 and ld.timed()[0] == 3


def test_reference_arm_runs_the_staged_reference():
    from baseline import vendor_ref
    vendor_ref.vendor()                      # no-op on the GPU box (no /root/reference): uses what travelled
    if vendor_ref.staged_dir() is None:
        pytest.skip("baseline/_ref not staged (no /root/reference in this environment)")
    import bench
    rate, ms, threads, n = bench.reference_step_rate(steps=2, warmup=1, batch=4, threads=2)
    assert n == 2 and rate > 0 and ms > 0
    trainer = sys.modules["trainer"]
    assert os.path.realpath(trainer.__file__).startswith(os.path.realpath(os.path.join(ROOT, "baseline", "_ref")))
    # staged files are byte-identical to the reference where it is present
    ref = "/root/reference/conditional_counteRGAN/mnist/trainer.py"
    if os.path.exists(ref):
        assert open(ref, "rb").read() == open(trainer.__file__, "rb").read()
