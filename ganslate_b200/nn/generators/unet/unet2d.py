"""Unet2D generator -- constructor, recursive module tree, parameter order and state_dict keys of
ganslate/nn/generators/unet/unet2d.py:17-157; forward runs on the fused sm_100a kernels.

A UnetSkipConnectionBlock is [LeakyReLU -> Conv4 s2 -> IN] down, the sub-block, [ReLU -> ConvT4 s2 -> IN (-> Dropout)]
up and `cat([x, model(x)], 1)` (unet2d.py:148-157).  The concatenation is not a kernel: the block allocates the
(Cx + outer_nc)-channel buffer once, the up-path InstanceNorm writes its channel slice and the skip tensor is copied
into the other; gradients of both consumers of x accumulate in its fp32 gradient buffer (nn/layers.py, Storage)."""
from dataclasses import dataclass

import torch
from torch import nn

from ganslate_b200 import configs, ops
from ganslate_b200._cabi import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_TANH
from ganslate_b200.nn import layers
from ganslate_b200.nn.utils import get_norm_layer_2d, get_norm_layer_3d, is_bias_before_norm


@dataclass
class Unet2DConfig(configs.base.BaseGeneratorConfig):
    num_downs: int = 7
    ngf: int = 64
    use_dropout: bool = False


class Unet2D(nn.Module):

    def __init__(self, in_channels, out_channels, num_downs, norm_type, ngf=64, use_dropout=False):
        super().__init__()
        # built from the innermost block outwards (unet2d.py:37-73)
        block = UnetSkipConnectionBlock(ngf * 8, ngf * 8, in_channels=None, submodule=None, norm_type=norm_type,
                                        innermost=True)
        for _ in range(num_downs - 5):
            block = UnetSkipConnectionBlock(ngf * 8, ngf * 8, in_channels=None, submodule=block, norm_type=norm_type,
                                            use_dropout=use_dropout)
        block = UnetSkipConnectionBlock(ngf * 4, ngf * 8, in_channels=None, submodule=block, norm_type=norm_type)
        block = UnetSkipConnectionBlock(ngf * 2, ngf * 4, in_channels=None, submodule=block, norm_type=norm_type)
        block = UnetSkipConnectionBlock(ngf, ngf * 2, in_channels=None, submodule=block, norm_type=norm_type)
        self.model = UnetSkipConnectionBlock(out_channels, ngf, in_channels=in_channels, submodule=block,
                                             outermost=True, norm_type=norm_type)

    def forward(self, input):
        params = list(self.parameters())
        ops._require_cuda(input, "network input")
        ops.ensure_packed(self)
        return layers.RunnerFn.apply(lambda tape, b0: (self.model.gb_run(tape, b0), ACT_TANH),
                                     (id(self), bool(self.training)), input, *params)


class Dropout(layers._Marker, nn.Dropout):
    """nn.Dropout(0.5) of the intermediate U-Net blocks (unet2d.py:143-144); applied by step_dropout."""


def step_dropout(tape, b, p):
    """In-place inverted dropout on the channel slice `b` (training mode only).  The mask is drawn with torch's
    Philox generator (CUDA-graph safe); backward scales the slice of the fp32 gradient by the same mask."""
    sl = b.st.t[..., b.c0:b.c0 + b.cw]
    mask = (torch.rand(sl.shape, device=sl.device) >= p).to(torch.bfloat16) * (1.0 / (1.0 - p))
    sl.mul_(mask)

    def bwd():
        if b.has_grad():
            b.st.grad[..., b.c0:b.c0 + b.cw].mul_(mask.float())

    if tape is not None:
        tape.steps.append(bwd)


class UnetSkipConnectionBlock(nn.Module):
    _dims = 2  # the 3-D generator (unet3d.py) is the same block over Conv3d / ConvTranspose3d / InstanceNorm3d

    def __init__(self, outer_nc, inner_nc, norm_type, in_channels=None, submodule=None, outermost=False,
                 innermost=False, use_dropout=False):
        super().__init__()
        self.outermost, self.innermost = outermost, innermost
        norm_layer = get_norm_layer_2d(norm_type) if self._dims == 2 else get_norm_layer_3d(norm_type)
        Conv = layers.Conv2d if self._dims == 2 else layers.Conv3d
        ConvT = layers.ConvTranspose2d if self._dims == 2 else layers.ConvTranspose3d
        use_bias = is_bias_before_norm(norm_type)
        if in_channels is None:
            in_channels = outer_nc
        self.in_nc, self.outer_nc = in_channels, outer_nc
        downconv = Conv(in_channels, inner_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
        downrelu = layers.LeakyReLU(0.2)
        downnorm = norm_layer(inner_nc)
        uprelu = layers.ReLU()
        upnorm = norm_layer(outer_nc)
        if outermost:
            upconv = ConvT(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1)
            model = [downconv] + [submodule] + [uprelu, upconv, layers.Tanh()]
        elif innermost:
            upconv = ConvT(inner_nc, outer_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
            model = [downrelu, downconv] + [uprelu, upconv, upnorm]
        else:
            upconv = ConvT(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
            model = [downrelu, downconv, downnorm] + [submodule] + [uprelu, upconv, upnorm]
            if use_dropout:
                model = model + [Dropout(0.5)]
        self.model = nn.Sequential(*model)
        # the layers by role (their indices in self.model differ per block kind); kept out of the module registry --
        # they already live in self.model and must not appear twice in state_dict()
        self.__dict__["_roles"] = dict(downconv=downconv, upconv=upconv, submodule=submodule,
                                       downnorm=None if (outermost or innermost) else downnorm,
                                       upnorm=None if outermost else upnorm)
        self._dropout = 0.5 if (use_dropout and not outermost and not innermost) else 0.0

    def gb_run(self, tape, x):
        """x: activation Buf (the outermost block gets the network input).  Returns the block output Buf:
        cat([x, model(x)]) for inner blocks, the raw output convolution (tanh applied on export) for the outermost."""
        R = self._roles
        sub, downconv, upconv = R["submodule"], R["downconv"], R["upconv"]
        if self.outermost:
            raw = layers.step_conv(tape, x, downconv)
            # two consumers (the sub-block's LeakyReLU and its skip copy): turn the raw convolution output into an
            # activation buffer whose fp32 gradient accumulates
            d = layers.step_norm_act(tape, raw, False, ACT_NONE, 0.0, 0, 1e-5)
            s = sub.gb_run(tape, d)
            r = layers.step_norm_act(tape, s, False, ACT_RELU, 0.0, 0, 1e-5)
            return layers.step_conv(tape, r, upconv)
        if x.channels % 8:
            raise NotImplementedError("Unet2D skip concatenation needs channel counts that are multiples of 8 (ngf % 8 == 0)")
        a = layers.step_norm_act(tape, x, False, ACT_LEAKY, 0.2, 0, 1e-5)
        if self.innermost:
            r = layers.step_conv(tape, a, downconv, ACT_RELU, 0.0)  # bias + ReLU in the epilogue
        else:
            raw = layers.step_conv(tape, a, downconv, want_stats=True)
            d = layers.step_norm_act(tape, raw, True, ACT_NONE, 0.0, 0, R["downnorm"].eps)
            s = sub.gb_run(tape, d)
            r = layers.step_norm_act(tape, s, False, ACT_RELU, 0.0, 0, 1e-5)
        raw_u = layers.step_conv(tape, r, upconv, want_stats=True)
        xcat = layers.new_like(x, x.channels + self.outer_nc)
        layers.step_norm_act(tape, x, False, ACT_NONE, 0.0, 0, 1e-5, out=xcat.slice(0, x.channels))
        up = xcat.slice(ops.pad8(x.channels), self.outer_nc)
        layers.step_norm_act(tape, raw_u, True, ACT_NONE, 0.0, 0, R["upnorm"].eps, out=up)
        if self._dropout > 0.0 and self.training:
            step_dropout(tape, up, self._dropout)
        return xcat

    def forward(self, x):
        raise RuntimeError("UnetSkipConnectionBlock is executed through Unet2D.forward")
