"""TEST INFRASTRUCTURE ONLY.  Makes the reference's own nn modules importable in the build container.

/root/reference is pure Python; its `nn` package imports with three import-only stubs (omegaconf, monai,
loguru) and a functional memcnn stand-in.  Nothing on the product path imports this;
it is used by oracle/make_golden.py and by the CPU tests that pin oracle/torch_oracle.py to the reference.
/root/reference does not exist on the GPU box -- callers must check `available()`.
"""
import os
import sys

REF_ROOT = os.environ.get("GANSLATE_REFERENCE", "/root/reference")
STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "ganslate"))


def setup():
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for p in (REF_ROOT, STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ganslate.configs.base  # noqa: F401  (modules reach configs.base as an attribute of the package)


def modules():
    """The reference classes on the hot path."""
    setup()
    from ganslate.nn.generators.resnet.resnet2d import Resnet2D
    from ganslate.nn.discriminators.patchgan.patchgan2d import PatchGAN2D
    from ganslate.nn.discriminators.patchgan.patchgan3d import PatchGAN3D
    from ganslate.nn.generators.unet.unet2d import Unet2D
    from ganslate.nn.generators.vnet.vnet3d import Vnet3D
    from ganslate.nn.losses.adversarial_loss import AdversarialLoss
    from ganslate.nn.losses.cyclegan_losses import CycleGANLosses
    from ganslate.nn.losses.pix2pix_losses import Pix2PixLoss
    from ganslate.nn.losses.cut_losses import PatchNCELoss
    from ganslate.nn.utils import init_weights
    from ganslate.data.utils.image_pool import ImagePool
    return dict(Resnet2D=Resnet2D, PatchGAN2D=PatchGAN2D, PatchGAN3D=PatchGAN3D, Unet2D=Unet2D, Vnet3D=Vnet3D,
                AdversarialLoss=AdversarialLoss, CycleGANLosses=CycleGANLosses, Pix2PixLoss=Pix2PixLoss,
                PatchNCELoss=PatchNCELoss, init_weights=init_weights, ImagePool=ImagePool)
