"""Resnet3D generator -- constructor, module list, parameter order and state_dict keys of
ganslate/nn/generators/resnet/resnet3d.py:14-91 (ReplicationPad3d instead of the 2-D version's ReflectionPad2d).

The 7x7x7 input / output convolutions have 343 taps, more than one launch of the implicit-GEMM kernels addresses
(GB_MAX_TAPS = 128): they run as seven (1, 7, 7) depth slabs accumulated in FP32 (ops.SlabConv).  Replicate padding is
a streaming copy (csrc/pad.cu)."""
from dataclasses import dataclass

from torch import nn

from ganslate_b200 import configs
from ganslate_b200.nn import layers
from ganslate_b200.nn.utils import get_norm_layer_3d, is_bias_before_norm


@dataclass
class Resnet3DConfig(configs.base.BaseGeneratorConfig):
    n_residual_blocks: int = 9


class Resnet3D(nn.Module):

    def __init__(self, in_channels, out_channels, norm_type, n_residual_blocks=9):
        super().__init__()
        norm_layer = get_norm_layer_3d(norm_type)
        use_bias = is_bias_before_norm(norm_type)
        model = [layers.ReplicationPad3d(3), layers.Conv3d(in_channels, 64, 7, bias=use_bias), norm_layer(64),
                 layers.ReLU(inplace=True)]
        in_features = 64
        out_features = in_features * 2
        for _ in range(2):
            model += [layers.Conv3d(in_features, out_features, 3, stride=2, padding=1, bias=use_bias),
                      norm_layer(out_features), layers.ReLU(inplace=True)]
            in_features = out_features
            out_features = in_features * 2
        for _ in range(n_residual_blocks):
            model += [ResidualBlock(in_features, norm_type)]
        out_features = in_features // 2
        for _ in range(2):
            model += [layers.ConvTranspose3d(in_features, out_features, 3, stride=2, padding=1, output_padding=1),
                      norm_layer(out_features), layers.ReLU(inplace=True)]
            in_features = out_features
            out_features = in_features // 2
        model += [layers.ReplicationPad3d(3), layers.Conv3d(64, out_channels, 7, bias=use_bias), layers.Tanh()]
        self.model = nn.Sequential(*model)

    def forward(self, x):
        return layers.run_network(self, list(self.model), x)


class ResidualBlock(nn.Module):

    def __init__(self, in_features, norm_type):
        super().__init__()
        norm_layer = get_norm_layer_3d(norm_type)
        use_bias = is_bias_before_norm(norm_type)
        conv_block = [layers.ReplicationPad3d(1), layers.Conv3d(in_features, in_features, 3, bias=use_bias),
                      norm_layer(in_features), layers.ReLU(inplace=True),
                      layers.ReplicationPad3d(1), layers.Conv3d(in_features, in_features, 3, bias=use_bias),
                      norm_layer(in_features)]
        self.conv_block = nn.Sequential(*conv_block)

    def gb_first_pad(self):
        return 0  # replicate padding is its own step, not a border of the producer's buffer

    def gb_run(self, tape, b, next_pad):
        # x + conv_block(x): the add is fused into the second InstanceNorm kernel
        return layers.run_sequence(tape, list(self.conv_block), b, final_pad=next_pad, residual=b)

    def forward(self, x):
        raise RuntimeError("ResidualBlock is executed through Resnet3D.forward / run_sequence")
