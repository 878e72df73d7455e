"""PatchGAN3D discriminator -- same constructor / module list / state_dict keys as
ganslate/nn/discriminators/patchgan/patchgan3d.py:17-65 (Conv3d k4, InstanceNorm3d, LeakyReLU 0.2); compute on the
same fused sm_100a kernels as the 2-D network (the implicit GEMM enumerates (n, z, y, x) output positions)."""
from dataclasses import dataclass
from typing import Tuple

from torch import nn

from ganslate_b200 import configs
from ganslate_b200.nn import layers
from ganslate_b200.nn.utils import get_norm_layer_3d, is_bias_before_norm


@dataclass
class PatchGAN3DConfig(configs.base.BaseDiscriminatorConfig):
    ndf: int = 64
    n_layers: int = 3
    kernel_size: Tuple[int] = (4, 4, 4)


class PatchGAN3D(nn.Module):

    def __init__(self, in_channels, ndf, n_layers, kernel_size, norm_type):
        super().__init__()
        norm_layer = get_norm_layer_3d(norm_type)
        use_bias = is_bias_before_norm(norm_type)
        kw, padw = tuple(kernel_size), 1
        seq = [layers.Conv3d(in_channels, ndf, kernel_size=kw, stride=2, padding=padw), layers.LeakyReLU(0.2, True)]
        mult = 1
        for n in range(1, n_layers):
            prev, mult = mult, min(2**n, 8)
            seq += [layers.Conv3d(ndf * prev, ndf * mult, kernel_size=kw, stride=2, padding=padw, bias=use_bias),
                    norm_layer(ndf * mult), layers.LeakyReLU(0.2, True)]
        prev, mult = mult, min(2**n_layers, 8)
        seq += [layers.Conv3d(ndf * prev, ndf * mult, kernel_size=kw, stride=1, padding=padw, bias=use_bias),
                norm_layer(ndf * mult), layers.LeakyReLU(0.2, True)]
        seq += [layers.Conv3d(ndf * mult, 1, kernel_size=kw, stride=1, padding=padw)]
        self.model = nn.Sequential(*seq)

    def forward(self, input):
        return layers.run_network(self, list(self.model), input)
