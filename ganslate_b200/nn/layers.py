"""Layer modules and the fused sequence interpreter.

The reference builds its networks as nn.Sequential lists of torch.nn layers (pad, conv, norm, activation;
e.g. ganslate/nn/generators/resnet/resnet2d.py:22-68).  To stay a drop-in -- identical `state_dict()` keys,
`parameters()` order and `init_weights` behaviour (ganslate/nn/utils.py:13-36 matches class names containing
"Conv") -- the layers here subclass the torch.nn classes for their parameter containers only.  Their compute is
never torch's: `run_sequence` walks the list and fuses [ReflectionPad] -> Conv -> [InstanceNorm] ->
[activation] -> [next ReflectionPad] into sm_100a kernel launches over channels-last bf16 buffers.
"""
from typing import List, Optional, Sequence

import torch
from torch import nn

from .. import ops
from .._cabi import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_TANH


def _t3(v):
    v = tuple(v) if not isinstance(v, int) else (v,)
    return (1,) * (3 - len(v)) + v if len(v) < 3 else v


def _p3(v):
    v = tuple(v) if not isinstance(v, int) else (v,)
    return (0,) * (3 - len(v)) + v if len(v) < 3 else v


class _ConvMixin:
    _transposed = False

    def conv_op(self) -> ops.ConvOp:
        op = self.__dict__.get("_gb_op")
        if op is None:
            if self.groups != 1 or any(d != 1 for d in self.dilation):
                raise NotImplementedError("ganslate_b200 convolutions support groups=1, dilation=1")
            if self.padding_mode != "zeros":
                raise NotImplementedError("use an explicit ReflectionPad module (as the reference networks do)")
            op = ops.ConvOp(self.in_channels, self.out_channels, _t3(self.kernel_size), _t3(self.stride),
                            _p3(self.padding), transposed=self._transposed,
                            output_padding=_p3(self.output_padding) if self._transposed else (0, 0, 0))
            self.__dict__["_gb_op"] = op
        return op

    def forward(self, x):  # noqa: D401
        raise RuntimeError("ganslate_b200 layers are executed by run_sequence() on CUDA buffers; "
                           "there is no eager/CPU path")


class Conv2d(_ConvMixin, nn.Conv2d):
    pass


class Conv3d(_ConvMixin, nn.Conv3d):
    pass


class ConvTranspose2d(_ConvMixin, nn.ConvTranspose2d):
    _transposed = True


class ConvTranspose3d(_ConvMixin, nn.ConvTranspose3d):
    _transposed = True


class _Marker:
    def forward(self, x):  # noqa: D401
        raise RuntimeError("ganslate_b200 layers are executed by run_sequence() on CUDA buffers; "
                           "there is no eager/CPU path")


class ReflectionPad2d(_Marker, nn.ReflectionPad2d):
    @property
    def pad_amount(self):
        p = self.padding
        if len(set(p)) != 1:
            raise NotImplementedError("asymmetric reflection padding")
        return int(p[0])


class InstanceNorm2d(_Marker, nn.InstanceNorm2d):
    pass


class InstanceNorm3d(_Marker, nn.InstanceNorm3d):
    pass


class ReLU(_Marker, nn.ReLU):
    pass


class LeakyReLU(_Marker, nn.LeakyReLU):
    pass


class Tanh(_Marker, nn.Tanh):
    pass


class Buf:
    """A channels-last bf16 buffer travelling through a network: tensor + its reflection border + logical channels."""
    __slots__ = ("t", "pad", "channels", "is_3d")

    def __init__(self, t, pad, channels, is_3d):
        self.t, self.pad, self.channels, self.is_3d = t, pad, channels, is_3d


def to_buf(x: torch.Tensor, pad: int) -> Buf:
    return Buf(ops.ToChannelsLastFn.apply(x, pad), pad, x.shape[1], x.dim() == 5)


def from_buf(b: Buf, act: int = ACT_NONE) -> torch.Tensor:
    if b.pad != 0:
        raise RuntimeError("cannot export a bordered buffer")
    return ops.FromChannelsLastFn.apply(b.t, b.channels, b.is_3d, act)


def run_network(mods: Sequence[nn.Module], x: torch.Tensor) -> torch.Tensor:
    """NC(D)HW fp32 in -> NC(D)HW fp32 out through the fused kernels. A trailing nn.Tanh (generator output,
    resnet2d.py:65) is evaluated in fp32 while the result is exported."""
    mods = flatten_modules(mods)
    act = ACT_NONE
    if len(mods) and isinstance(mods[-1], Tanh):
        mods, act = mods[:-1], ACT_TANH
    b = to_buf(x, first_pad(mods))
    b = run_sequence(mods, b)
    return from_buf(b, act)


def _is_norm(m):
    return isinstance(m, (InstanceNorm2d, InstanceNorm3d))


def _act_of(m):
    if isinstance(m, ReLU):
        return ACT_RELU, 0.0
    if isinstance(m, LeakyReLU):
        return ACT_LEAKY, float(m.negative_slope)
    if isinstance(m, Tanh):
        return ACT_TANH, 0.0
    return None


def first_pad(mods: Sequence[nn.Module]) -> int:
    """Reflection border the first layer of `mods` wants on its input (0 if it does not start with a pad)."""
    if len(mods) == 0:
        return 0
    m = mods[0]
    if isinstance(m, ReflectionPad2d):
        return m.pad_amount
    if hasattr(m, "gb_first_pad"):
        return m.gb_first_pad()
    return 0


def flatten_modules(mods) -> List[nn.Module]:
    out = []
    for m in mods:
        if isinstance(m, nn.Sequential):
            out += flatten_modules(list(m))
        else:
            out.append(m)
    return out


def run_sequence(mods: Sequence[nn.Module], b: Buf, final_pad: int = 0, residual: Optional[Buf] = None) -> Buf:
    """Execute a list of layers on buffer `b`, fusing pad/conv/norm/activation groups.

    final_pad: reflection border wanted on the last produced buffer (what follows this sequence).
    residual:  added to the output of the LAST norm group (ResidualBlock: x + conv_block(x)).
    """
    mods = flatten_modules(mods)
    i, n = 0, len(mods)
    pending_pad = 0  # border announced by a ReflectionPad module for the next convolution
    while i < n:
        m = mods[i]
        if isinstance(m, ReflectionPad2d):
            if b.pad != m.pad_amount:
                if b.pad != 0:
                    raise RuntimeError("buffer already carries a different reflection border")
                # producer did not materialise the border: copy-with-border pass
                t = ops.NormActFn.apply(b.t, None, False, ACT_NONE, 0.0, m.pad_amount, 0, 1e-5)
                b = Buf(t, m.pad_amount, b.channels, b.is_3d)
            i += 1
            pending_pad = m.pad_amount
            if i >= n or not isinstance(mods[i], _ConvMixin):
                raise RuntimeError("ReflectionPad must be followed by a convolution")
            continue
        if isinstance(m, _ConvMixin):
            op = m.conv_op()
            if b.pad != pending_pad:
                raise RuntimeError(f"convolution input carries border {b.pad}, expected {pending_pad}")
            pending_pad = 0
            j = i + 1
            norm = j < n and _is_norm(mods[j])
            if norm:
                nm = mods[j]
                if nm.affine or nm.track_running_stats:
                    raise NotImplementedError("InstanceNorm with affine/running stats (the reference uses neither)")
                j += 1
            act = _act_of(mods[j]) if j < n else None
            if act is not None:
                j += 1
            act_id, slope = act if act is not None else (ACT_NONE, 0.0)
            # border wanted by whatever consumes this group's output
            nxt = first_pad(mods[j:]) if j < n else final_pad
            is_last_group = j >= n
            res = residual if (is_last_group and residual is not None) else None
            x = b.t  # the whole (bordered) buffer is the convolution input; the border replaces ReflectionPad
            if norm or res is not None or nxt > 0:
                raw = ops.ConvFn.apply(x, m.weight, m.bias, op, ACT_NONE, 0.0)
                eps = mods[i + 1].eps if norm else 1e-5
                t = ops.NormActFn.apply(raw, res.t if res is not None else None, norm, act_id, slope, nxt,
                                        res.pad if res is not None else 0, eps)
                b = Buf(t, nxt, m.out_channels, b.is_3d)
            else:
                # no normalisation: bias + activation run in the convolution epilogue
                t = ops.ConvFn.apply(x, m.weight, m.bias, op, act_id, slope)
                b = Buf(t, 0, m.out_channels, b.is_3d)
            i = j
            continue
        if hasattr(m, "gb_run"):
            nxt = first_pad(mods[i + 1:]) if i + 1 < n else final_pad
            b = m.gb_run(b, nxt)
            i += 1
            continue
        if isinstance(m, (nn.Identity,)):
            i += 1
            continue
        raise NotImplementedError(f"run_sequence: unsupported layer {type(m).__name__} at position {i}")
    if residual is not None and not any(isinstance(m, _ConvMixin) for m in mods):
        raise RuntimeError("residual requested on a sequence without convolutions")
    return b
