from .pix2pix import Pix2PixConditionalGAN  # noqa: F401
