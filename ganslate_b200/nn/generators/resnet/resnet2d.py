"""Resnet2D generator -- same constructor, module list, parameter order and state_dict keys as
ganslate/nn/generators/resnet/resnet2d.py:14-93; compute runs on the fused sm_100a kernels."""
from dataclasses import dataclass

from torch import nn

from ganslate_b200 import configs
from ganslate_b200.nn import layers
from ganslate_b200.nn.utils import get_norm_layer_2d, is_bias_before_norm


@dataclass
class Resnet2DConfig(configs.base.BaseGeneratorConfig):
    n_residual_blocks: int = 9


class Resnet2D(nn.Module):

    def __init__(self, in_channels, out_channels, norm_type, n_residual_blocks=9):
        super().__init__()
        norm_layer = get_norm_layer_2d(norm_type)
        use_bias = is_bias_before_norm(norm_type)

        model = [
            layers.ReflectionPad2d(3),
            layers.Conv2d(in_channels, 64, 7, bias=use_bias),
            norm_layer(64),
            layers.ReLU(inplace=True)
        ]
        in_features = 64
        out_features = in_features * 2
        for _ in range(2):
            model += [
                layers.Conv2d(in_features, out_features, 3, stride=2, padding=1, bias=use_bias),
                norm_layer(out_features),
                layers.ReLU(inplace=True)
            ]
            in_features = out_features
            out_features = in_features * 2

        for _ in range(n_residual_blocks):
            model += [ResidualBlock(in_features, norm_type)]

        # reference: `self.encoder` aliases the first 19 modules (resnet2d.py:46) -- same objects, so the
        # state_dict exposes them under both `encoder.*` and `model.*`
        self.encoder = nn.ModuleList(model)

        out_features = in_features // 2
        for _ in range(2):
            model += [
                layers.ConvTranspose2d(in_features, out_features, 3, stride=2, padding=1, output_padding=1),
                norm_layer(out_features),
                layers.ReLU(inplace=True)
            ]
            in_features = out_features
            out_features = in_features // 2

        model += [layers.ReflectionPad2d(3), layers.Conv2d(64, out_channels, 7, bias=use_bias), layers.Tanh()]
        self.model = nn.Sequential(*model)

    def forward(self, x):
        return layers.run_network(self, list(self.model), x)


class ResidualBlock(nn.Module):

    def __init__(self, in_features, norm_type):
        super().__init__()
        norm_layer = get_norm_layer_2d(norm_type)
        use_bias = is_bias_before_norm(norm_type)
        conv_block = [
            layers.ReflectionPad2d(1),
            layers.Conv2d(in_features, in_features, 3, bias=use_bias),
            norm_layer(in_features),
            layers.ReLU(inplace=True),
            layers.ReflectionPad2d(1),
            layers.Conv2d(in_features, in_features, 3, bias=use_bias),
            norm_layer(in_features)
        ]
        self.conv_block = nn.Sequential(*conv_block)

    # hooks for layers.run_sequence
    def gb_first_pad(self):
        return layers.first_pad(list(self.conv_block))

    def gb_run(self, tape, b, next_pad):
        # x + conv_block(x): the add is fused into the second InstanceNorm kernel
        return layers.run_sequence(tape, list(self.conv_block), b, final_pad=next_pad, residual=b)

    def forward(self, x):
        raise RuntimeError("ResidualBlock is executed through Resnet2D.forward / run_sequence")
