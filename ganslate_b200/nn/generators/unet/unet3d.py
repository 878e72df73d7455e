"""Unet3D generator -- ganslate/nn/generators/unet/unet3d.py:17-157: the Unet2D recursion over Conv3d /
ConvTranspose3d (k 4, s 2, p 1) / InstanceNorm3d.  Same kernels, same tape steps as unet2d.py."""
from dataclasses import dataclass

from torch import nn

from ganslate_b200 import configs, ops
from ganslate_b200._cabi import ACT_TANH
from ganslate_b200.nn import layers
from ganslate_b200.nn.generators.unet import unet2d


@dataclass
class Unet3DConfig(configs.base.BaseGeneratorConfig):
    num_downs: int = 7
    ngf: int = 64
    use_dropout: bool = False


class UnetSkipConnectionBlock(unet2d.UnetSkipConnectionBlock):
    _dims = 3


class Unet3D(nn.Module):

    def __init__(self, in_channels, out_channels, num_downs, norm_type, ngf=64, use_dropout=False):
        super().__init__()
        B = UnetSkipConnectionBlock
        block = B(ngf * 8, ngf * 8, in_channels=None, submodule=None, norm_type=norm_type, innermost=True)
        for _ in range(num_downs - 5):
            block = B(ngf * 8, ngf * 8, in_channels=None, submodule=block, norm_type=norm_type, use_dropout=use_dropout)
        block = B(ngf * 4, ngf * 8, in_channels=None, submodule=block, norm_type=norm_type)
        block = B(ngf * 2, ngf * 4, in_channels=None, submodule=block, norm_type=norm_type)
        block = B(ngf, ngf * 2, in_channels=None, submodule=block, norm_type=norm_type)
        self.model = B(out_channels, ngf, in_channels=in_channels, submodule=block, outermost=True, norm_type=norm_type)

    def forward(self, input):
        params = list(self.parameters())
        ops._require_cuda(input, "network input")
        ops.ensure_packed(self)
        return layers.RunnerFn.apply(lambda tape, b0: (self.model.gb_run(tape, b0), ACT_TANH),
                                     (id(self), bool(self.training)), input, *params)
