from .cyclegan import CycleGAN  # noqa: F401
from .cut import CUT  # noqa: F401
from .revgan import RevGAN  # noqa: F401
