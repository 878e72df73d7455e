from .resnet.resnet2d import Resnet2D  # noqa: F401
