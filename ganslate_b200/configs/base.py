"""Config dataclasses of the hot path, mirroring ganslate/configs/base.py:19-62.

The reference uses OmegaConf structured configs; OmegaConf is optional here (absent in this image), the
dataclasses are plain and use default_factory so they import on Python >= 3.11 (the reference's
`field: X = X()` defaults do not, ganslate/configs/training.py:35).
"""
from dataclasses import dataclass, field
from typing import Optional, Tuple

MISSING = "???"


@dataclass
class BaseOptimizerConfig:
    adversarial_loss_type: str = "lsgan"
    beta1: float = 0.5
    beta2: float = 0.999
    lr_D: float = 0.0001
    lr_G: float = 0.0002


@dataclass
class GeneratorInOutChannelsConfig:
    AB: Tuple[int, int] = MISSING
    BA: Optional[Tuple[int, int]] = None  # reference: interpolates to AB


@dataclass
class BaseGeneratorConfig:
    _target_: str = MISSING
    in_out_channels: GeneratorInOutChannelsConfig = field(default_factory=GeneratorInOutChannelsConfig)


@dataclass
class DiscriminatorInChannelsConfig:
    B: int = MISSING
    A: Optional[int] = None  # reference: interpolates to B


@dataclass
class BaseDiscriminatorConfig:
    _target_: str = MISSING
    in_channels: DiscriminatorInChannelsConfig = field(default_factory=DiscriminatorInChannelsConfig)


@dataclass
class BaseGANConfig:
    _target_: str = MISSING
    norm_type: str = "instance"
    weight_init_type: str = "normal"
    weight_init_gain: float = 0.02
    optimizer: BaseOptimizerConfig = field(default_factory=BaseOptimizerConfig)
    generator: BaseGeneratorConfig = field(default_factory=BaseGeneratorConfig)
    discriminator: Optional[BaseDiscriminatorConfig] = None
