"""Drop-in mirror of the reference's ``conditional_counteRGAN/mnist`` hot path.

    from pcg_b200.mnist.models.generator import ResidualGenerator      # models/generator.py:25
    from pcg_b200.mnist.models.discriminator import Discriminator      # models/discriminator.py:5
    from pcg_b200.mnist.models.classifier import CNNClassifier         # models/classifier.py:4
    from pcg_b200.mnist.trainer import train_countergan, build_mask    # trainer.py:76, :45
"""
from .plan import MnistStepPlan, StepConfig  # noqa: F401
