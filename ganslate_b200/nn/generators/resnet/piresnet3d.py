"""Piresnet3D (partially-invertible ResNet, the generator of the shipped BraTS RevGAN config) -- constructor, module
tree, parameter order and state_dict keys of ganslate/nn/generators/resnet/piresnet3d.py:28-119; forward(x,
inverse=False) runs on the fused sm_100a kernels.

downconv: ReplicationPad3d(2) -> Conv3d 5^3 -> IN -> ReLU -> Conv3d 3^3 s2 p1 -> IN -> ReLU
core:     `depth` additive couplings of [IN, ReplicationPad3d(1), Conv3d 3^3, IN, ReLU] on channel halves
upconv:   ConvTranspose3d 3^3 s2 p1 op1 -> IN -> ReLU -> ReplicationPad3d(2) -> Conv3d 5^3 -> Tanh

Replicate padding is materialised by a streaming copy (csrc/pad.cu; a TMA box cannot clamp coordinates); the
couplings work on channel-slice views like V-Net's (nn/invertible.py)."""
from dataclasses import dataclass

from torch import nn

from ganslate_b200 import configs, ops
from ganslate_b200._cabi import ACT_TANH
from ganslate_b200.configs.base import MISSING
from ganslate_b200.nn import invertible, layers
from ganslate_b200.nn.utils import get_norm_layer_3d, is_bias_before_norm


@dataclass
class Piresnet3DConfig(configs.base.BaseGeneratorConfig):
    use_memory_saving: bool = True
    use_inverse: bool = True
    first_layer_channels: int = 32
    depth: int = MISSING


class Piresnet3D(nn.Module):

    def __init__(self, in_channels, out_channels, norm_type, depth, first_layer_channels=64, use_memory_saving=True,
                 use_inverse=True):
        super().__init__()
        keep_input = not use_memory_saving
        norm_layer = get_norm_layer_3d(norm_type)
        use_bias = is_bias_before_norm(norm_type)
        self.use_inverse = use_inverse

        self.downconv_ab = self.build_downconv(in_channels, norm_layer, first_layer_channels, use_bias)
        self.upconv_ab = self.build_upconv(out_channels, norm_layer, first_layer_channels, use_bias)
        if use_inverse:
            self.downconv_ba = self.build_downconv(in_channels, norm_layer, first_layer_channels, use_bias)
            self.upconv_ba = self.build_upconv(out_channels, norm_layer, first_layer_channels, use_bias)

        inv_block = _base_inv_block(first_layer_channels * 2, norm_layer, use_bias)
        self.core = invertible.InvertibleSequence(inv_block, depth, keep_input)

    @staticmethod
    def build_downconv(in_channels, norm_layer, c, use_bias):
        return nn.Sequential(layers.ReplicationPad3d(2),
                             layers.Conv3d(in_channels, c, kernel_size=5, stride=1, padding=0, bias=use_bias),
                             norm_layer(c), layers.ReLU(inplace=True),
                             layers.Conv3d(c, c * 2, kernel_size=3, stride=2, padding=1, bias=use_bias),
                             norm_layer(c * 2), layers.ReLU(inplace=True))

    @staticmethod
    def build_upconv(out_channels, norm_layer, c, use_bias):
        return nn.Sequential(layers.ConvTranspose3d(c * 2, c, kernel_size=3, stride=2, padding=1, output_padding=1,
                                                    bias=use_bias),
                             norm_layer(c), layers.ReLU(inplace=True), layers.ReplicationPad3d(2),
                             layers.Conv3d(c, out_channels, kernel_size=5, padding=0), layers.Tanh())

    def _run(self, tape, b0, inverse):
        down, up = (self.downconv_ba, self.upconv_ba) if inverse else (self.downconv_ab, self.upconv_ab)
        b = layers.run_sequence(tape, list(down), b0)
        b = self.core.gb_run_coupling(tape, b, inverse)
        b = layers.run_sequence(tape, list(up)[:-1], b)  # the trailing Tanh is evaluated in fp32 while exporting
        return b, ACT_TANH

    def forward(self, x, inverse=False):
        if inverse and not self.use_inverse:
            raise ValueError("Trying to perform inverse forward while `use_inverse` flag is turned off.")
        params = list(self.parameters())
        ops._require_cuda(x, "network input")
        ops.ensure_packed(self)
        return layers.RunnerFn.apply(lambda tape, b0: self._run(tape, b0, inverse), (id(self), bool(inverse)), x, *params)


def _base_inv_block(n_channels, norm_layer, use_bias):
    n_channels = n_channels // 2  # the coupling works on channel halves
    return nn.Sequential(norm_layer(n_channels), layers.ReplicationPad3d(1),
                         layers.Conv3d(n_channels, n_channels, kernel_size=3, padding=0, bias=use_bias),
                         norm_layer(n_channels), layers.ReLU(inplace=True))
