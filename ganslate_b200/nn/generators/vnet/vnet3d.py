"""Partially-invertible V-Net 3D generator -- constructor, module tree, parameter order and state_dict keys of
ganslate/nn/generators/vnet/vnet3d.py:27-267; forward(x, inverse=False) runs on the fused sm_100a kernels.

Channel splits, skip concatenations and the additive couplings are channel-slice VIEWS of channels-last buffers:
a coupling branch reads its half through a strided view and its fused norm/PReLU kernel writes `x1 + F(x2)`
straight into the output half; an up-block's transposed convolution writes the first half of the concatenation
buffer and the skip tensor is copied into the second."""
from dataclasses import dataclass
from typing import Tuple

import torch
from torch import nn

from ganslate_b200 import configs, ops
from ganslate_b200._cabi import ACT_NONE, ACT_PRELU, ACT_TANH
from ganslate_b200.nn import invertible, layers
from ganslate_b200.nn.utils import (get_conv_layer_3d, get_conv_transpose_layer_3d, get_norm_layer_3d,
                                    is_bias_before_norm)


@dataclass
class Vnet3DConfig(configs.base.BaseGeneratorConfig):
    use_memory_saving: bool = False
    use_inverse: bool = False
    first_layer_channels: int = 16
    down_blocks: Tuple[int] = (1, 2, 3, 2)
    up_blocks: Tuple[int] = (2, 2, 1, 1)
    is_separable: bool = False


def _conv_norm_prelu(tape, b, seq, out=None):
    """[conv, norm, PReLU] group."""
    raw = layers.step_conv_any(tape, b, seq[0], want_stats=True)
    return layers.step_norm_act(tape, raw, True, ACT_PRELU, 0.0, 0, seq[1].eps, prelu=seq[2], out=out)


class Vnet3D(nn.Module):

    def __init__(self, in_channels, out_channels, norm_type, first_layer_channels=16, down_blocks=(1, 2, 3, 2),
                 up_blocks=(2, 2, 1, 1), use_memory_saving=True, use_inverse=True, is_separable=False):
        super().__init__()
        disable_invertibles = use_memory_saving is False and use_inverse is False
        if first_layer_channels % in_channels:
            raise ValueError("`first_layer_channels` has to be divisible by `in_channels`.")
        if len(down_blocks) != len(up_blocks):
            raise ValueError("Number of `down_blocks` and `up_blocks` has to be equal.")
        keep_input = not use_memory_saving
        norm_layer = get_norm_layer_3d(norm_type)
        use_bias = is_bias_before_norm(norm_type)
        self.use_inverse = use_inverse
        c0 = first_layer_channels

        self.in_ab = InputBlock(in_channels, c0, norm_layer, use_bias, is_separable)
        if use_inverse:
            self.in_ba = InputBlock(in_channels, c0, norm_layer, use_bias, is_separable)
        self.out_ab = OutBlock(c0 * 2, out_channels, norm_layer, use_bias, is_separable)
        if use_inverse:
            self.out_ba = OutBlock(c0 * 2, out_channels, norm_layer, use_bias, is_separable)

        downs, factors = [], []
        for i, num_convs in enumerate(down_blocks):
            factor = 2**i
            downs.append(DownBlock(c0 * factor, num_convs, norm_layer, use_bias, keep_input, use_inverse,
                                   disable_invertibles, is_separable))
            factors.append(factor)
        self.downs = nn.ModuleList(downs)
        self.encoder = nn.ModuleList([self.in_ab]).extend(self.downs)  # vnet3d.py:89 (same module objects)

        up_factors = [f * 2 for f in reversed(factors)]
        ups = [UpBlock(c0 * up_factors[0], c0 * up_factors[0], up_blocks[0], norm_layer, use_bias, keep_input,
                       use_inverse, disable_invertibles, is_separable)]
        for i, num_convs in enumerate(up_blocks[1:]):
            ups.append(UpBlock(c0 * up_factors[i], c0 * up_factors[i + 1], num_convs, norm_layer, use_bias, keep_input,
                               use_inverse, disable_invertibles, is_separable))
        self.ups = nn.ModuleList(ups)

    def _run(self, tape, b0, inverse):
        in_block, out_block = (self.in_ba, self.out_ba) if inverse else (self.in_ab, self.out_ab)
        out1 = in_block.gb_run(tape, b0)
        down_outs = []
        for i, down in enumerate(self.downs):
            down_outs.append(down.gb_run(tape, out1 if i == 0 else down_outs[-1], inverse))
        rev = list(reversed(down_outs))
        out = rev[0]
        for i, up in enumerate(self.ups):
            skip = out1 if i == len(self.ups) - 1 else rev[i + 1]
            out = up.gb_run(tape, out, skip, inverse)
        return out_block.gb_run(tape, out), ACT_TANH

    def forward(self, x, inverse=False):
        if inverse and not self.use_inverse:
            raise ValueError("Trying to perform inverse forward while `use_inverse` flag is turned off.")
        params = list(self.parameters())
        ops._require_cuda(x, "network input")
        ops.ensure_packed(self)
        return layers.RunnerFn.apply(lambda tape, b0: self._run(tape, b0, inverse), (id(self), bool(inverse)), x, *params)


class InputBlock(nn.Module):

    def __init__(self, in_channels, out_channels, norm_layer, use_bias, is_separable=False):
        super().__init__()
        self.n_repeats = out_channels // in_channels
        self.conv1 = get_conv_layer_3d(is_separable)(in_channels, out_channels, kernel_size=5, padding=2, bias=use_bias)
        self.bn1 = norm_layer(out_channels)
        self.relu = layers.PReLU(out_channels)

    def gb_run(self, tape, b):
        # PReLU(IN(conv(x)) + x repeated over channels)
        raw = layers.step_conv_any(tape, b, self.conv1, want_stats=True)
        rep = layers.step_channel_repeat(tape, b, self.n_repeats)
        return layers.step_norm_act(tape, raw, True, ACT_PRELU, 0.0, 0, self.bn1.eps, residual=rep, prelu=self.relu,
                                    res_before_act=True)


class DownBlock(nn.Module):

    def __init__(self, in_channels, n_conv_blocks, norm_layer, use_bias, keep_input, use_inverse, disable_invertibles,
                 is_separable=False):
        super().__init__()
        self.is_separable = is_separable
        out_channels = 2 * in_channels
        self.down_conv_ab = self.build_down_conv(in_channels, out_channels, norm_layer, use_bias)
        if use_inverse:
            self.down_conv_ba = self.build_down_conv(in_channels, out_channels, norm_layer, use_bias)
        inv_block = _base_inv_block(out_channels, norm_layer, use_bias, is_separable)
        self.core = invertible.InvertibleSequence(inv_block, n_conv_blocks, keep_input, disable_invertibles)
        self.relu = layers.PReLU(out_channels)

    def build_down_conv(self, in_channels, out_channels, norm_layer, use_bias):
        conv_layer = get_conv_layer_3d(self.is_separable)
        return nn.Sequential(conv_layer(in_channels, out_channels, kernel_size=2, stride=2, bias=use_bias),
                             norm_layer(out_channels), layers.PReLU(out_channels))

    def gb_run(self, tape, b, inverse=False):
        down = _conv_norm_prelu(tape, b, self.down_conv_ba if inverse else self.down_conv_ab)
        out = self.core.gb_run_coupling(tape, down, inverse)
        # PReLU(out + down)
        return layers.step_norm_act(tape, out, False, ACT_PRELU, 0.0, 0, 1e-5, residual=down, prelu=self.relu,
                                    res_before_act=True)


class UpBlock(nn.Module):

    def __init__(self, in_channels, out_channels, n_conv_blocks, norm_layer, use_bias, keep_input, use_inverse,
                 disable_invertibles, is_separable=False):
        super().__init__()
        self.is_separable = is_separable
        self.out_channels = out_channels
        self.up_conv_ab = self.build_up_conv(in_channels, out_channels, norm_layer, use_bias)
        if use_inverse:
            self.up_conv_ba = self.build_up_conv(in_channels, out_channels, norm_layer, use_bias)
        inv_block = _base_inv_block(out_channels, norm_layer, use_bias, is_separable)
        self.core = invertible.InvertibleSequence(inv_block, n_conv_blocks, keep_input, disable_invertibles)
        self.relu = layers.PReLU(out_channels)

    def build_up_conv(self, in_channels, out_channels, norm_layer, use_bias):
        conv_transp_layer = get_conv_transpose_layer_3d(self.is_separable)
        return nn.Sequential(conv_transp_layer(in_channels, out_channels // 2, kernel_size=2, stride=2, bias=use_bias),
                             norm_layer(out_channels // 2), layers.PReLU(out_channels // 2))

    def gb_run(self, tape, b, skip, inverse=False):
        seq = self.up_conv_ba if inverse else self.up_conv_ab
        half = self.out_channels // 2
        raw = layers.step_conv_any(tape, b, seq[0], want_stats=True)
        # torch.cat((up, skipx), 1) without a cat kernel: both halves are written into one buffer
        N, D, H, W, _ = raw.t.shape
        xcat = layers.Buf(torch.empty((N, D, H, W, self.out_channels), dtype=raw.t.dtype, device=raw.t.device), 0,
                          self.out_channels, raw.is_3d)
        layers.step_norm_act(tape, raw, True, ACT_PRELU, 0.0, 0, seq[1].eps, prelu=seq[2], out=xcat.slice(0, half))
        layers.step_norm_act(tape, skip, False, ACT_NONE, 0.0, 0, 1e-5, out=xcat.slice(half, half))
        out = self.core.gb_run_coupling(tape, xcat, inverse)
        return layers.step_norm_act(tape, out, False, ACT_PRELU, 0.0, 0, 1e-5, residual=xcat, prelu=self.relu,
                                    res_before_act=True)


class OutBlock(nn.Module):

    def __init__(self, in_channels, out_channels, norm_layer, use_bias, is_separable=False):
        super().__init__()
        conv_layer = get_conv_layer_3d(is_separable)
        self.conv1 = conv_layer(in_channels, in_channels, kernel_size=5, padding=2, bias=use_bias)
        self.bn1 = norm_layer(in_channels)
        self.relu1 = layers.PReLU(in_channels)
        self.conv2 = conv_layer(in_channels, out_channels, kernel_size=1)
        self.tanh = layers.Tanh()

    def gb_run(self, tape, b):
        raw = layers.step_conv_any(tape, b, self.conv1, want_stats=True)
        a = layers.step_norm_act(tape, raw, True, ACT_PRELU, 0.0, 0, self.bn1.eps, prelu=self.relu1)
        return layers.step_conv_any(tape, a, self.conv2)  # tanh is evaluated in fp32 while exporting


def _base_inv_block(n_channels, norm_layer, use_bias, is_separable=False):
    n_channels = n_channels // 2  # the coupling works on channel halves
    return nn.Sequential(get_conv_layer_3d(is_separable)(n_channels, n_channels, kernel_size=5, padding=2, bias=use_bias),
                         norm_layer(n_channels), layers.PReLU(n_channels))
