from .patchgan.patchgan2d import PatchGAN2D  # noqa: F401
from .patchgan.patchgan3d import PatchGAN3D  # noqa: F401
