"""PatchGAN2D discriminator -- same constructor / module list / state_dict keys as
ganslate/nn/discriminators/patchgan/patchgan2d.py:17-66; compute on the fused sm_100a kernels."""
from dataclasses import dataclass
from typing import Tuple

from torch import nn

from ganslate_b200 import configs
from ganslate_b200.nn import layers
from ganslate_b200.nn.utils import get_norm_layer_2d, is_bias_before_norm


@dataclass
class PatchGAN2DConfig(configs.base.BaseDiscriminatorConfig):
    ndf: int = 64
    n_layers: int = 3
    kernel_size: Tuple[int] = (4, 4)


class PatchGAN2D(nn.Module):

    def __init__(self, in_channels, ndf, n_layers, kernel_size, norm_type):
        super().__init__()
        norm_layer = get_norm_layer_2d(norm_type)
        use_bias = is_bias_before_norm(norm_type)
        kw = tuple(kernel_size)
        padw = 1
        sequence = [layers.Conv2d(in_channels, ndf, kernel_size=kw, stride=2, padding=padw), layers.LeakyReLU(0.2, True)]
        nf_mult = 1
        for n in range(1, n_layers):
            nf_mult_prev = nf_mult
            nf_mult = min(2**n, 8)
            sequence += [
                layers.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=2, padding=padw, bias=use_bias),
                norm_layer(ndf * nf_mult),
                layers.LeakyReLU(0.2, True)
            ]
        nf_mult_prev = nf_mult
        nf_mult = min(2**n_layers, 8)
        sequence += [
            layers.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=1, padding=padw, bias=use_bias),
            norm_layer(ndf * nf_mult),
            layers.LeakyReLU(0.2, True)
        ]
        sequence += [layers.Conv2d(ndf * nf_mult, 1, kernel_size=kw, stride=1, padding=padw)]
        self.model = nn.Sequential(*sequence)

    def forward(self, input):
        return layers.run_network(self, list(self.model), input)
