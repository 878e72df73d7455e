from .resnet.resnet2d import Resnet2D  # noqa: F401
from .resnet.piresnet3d import Piresnet3D  # noqa: F401
from .resnet.resnet3d import Resnet3D  # noqa: F401
from .unet.unet2d import Unet2D  # noqa: F401
from .unet.unet3d import Unet3D  # noqa: F401
from .vnet.vnet3d import Vnet3D  # noqa: F401
