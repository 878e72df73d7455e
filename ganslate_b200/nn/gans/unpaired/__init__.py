from .cyclegan import CycleGAN  # noqa: F401
