from .patchgan.patchgan2d import PatchGAN2D  # noqa: F401
