"""Recipe that stages the UNMODIFIED reference files of the hot path under ``baseline/_ref/`` (git-ignored, but shipped
to the GPU box with the repository snapshot), so that ``bench.py --impl reference`` can drive the reference's own
``train_countergan`` on the box's host cores, where ``/root/reference`` does not exist.

Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present (the build container).  Only the files the
MNIST CounteRGAN iteration imports are staged (trainer + the three model definitions); ``config.py`` is NOT (it carries a
hard-coded API key, config.py:29, and the bench supplies its own hyper-parameter namespace with the values of
config.py:4-15).  Nothing under ``baseline/_ref/`` is ever committed or imported by the product path.
"""
import filecmp
import os
import shutil

REF_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = [
    "conditional_counteRGAN/mnist/trainer.py",
    "conditional_counteRGAN/mnist/models/generator.py",
    "conditional_counteRGAN/mnist/models/discriminator.py",
    "conditional_counteRGAN/mnist/models/classifier.py",
]


def vendor():
    """Copies FILES byte for byte; returns the list of staged paths ([] when the reference is absent)."""
    if not os.path.isdir(os.path.join(REF_ROOT, "conditional_counteRGAN")):
        return []
    out = []
    for rel in FILES:
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
        out.append(dst)
    return out


def staged_dir():
    """Directory to put first on sys.path (the reference uses bare module names), or None when nothing is staged."""
    d = os.path.join(DST, "conditional_counteRGAN", "mnist")
    return d if all(os.path.exists(os.path.join(DST, f)) for f in FILES) else None


if __name__ == "__main__":
    print("\n".join(vendor()) or "reference not present: nothing staged")
