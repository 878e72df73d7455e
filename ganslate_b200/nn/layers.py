"""Layer modules and the fused sequence interpreter.

The reference builds its networks as nn.Sequential lists of torch.nn layers (pad, conv, norm, activation;
e.g. ganslate/nn/generators/resnet/resnet2d.py:22-68).  To stay a drop-in -- identical `state_dict()` keys,
`parameters()` order and `init_weights` behaviour (ganslate/nn/utils.py:13-36 matches class names containing
"Conv") -- the layers here subclass the torch.nn classes for their parameter containers only.  Their compute is
never torch's: `run_sequence` walks the list and fuses [ReflectionPad] -> Conv -> [InstanceNorm] ->
[activation] -> [next ReflectionPad] into sm_100a kernel launches over channels-last bf16 buffers.

A whole network is ONE torch.autograd.Function (`NetworkFn`): the forward pass records a tape of fused steps,
the backward pass replays it in reverse.  Keeping the intermediate gradients out of autograd lets activation
gradients stay fp32 (autograd would cast them to the bf16 dtype of the forward buffers) and lets residual
branches accumulate into one gradient buffer inside the dgrad epilogue instead of through extra add kernels.
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
    """A channels-last bf16 buffer travelling through a network.

    t: tensor (N, D, H+2*pad, W+2*pad, Cpad); pad: materialised reflection border; channels: logical channels;
    raw: True if t is a bare convolution output (its gradient is the bf16 MMA operand of dgrad/wgrad),
    False for an activation buffer (fp32 gradient)."""
    __slots__ = ("t", "pad", "channels", "is_3d", "raw", "grad", "needs_grad_flag", "want_dbias", "dbias")

    def __init__(self, t, pad, channels, is_3d, raw=False):
        self.t, self.pad, self.channels, self.is_3d, self.raw = t, pad, channels, is_3d, raw
        self.grad = None  # filled during Tape.backward
        self.needs_grad_flag = True  # False only for a network input that does not require grad
        self.want_dbias = False      # the producing conv's bias needs a gradient (fused into the consumer's backward)
        self.dbias = None


class Tape:
    """Reverse-mode tape of fused steps; every step is a closure that consumes the gradient of its output
    Buf(s) and produces / accumulates the gradients of its inputs."""

    def __init__(self, param_needs_grad, input_needs_grad):
        self.steps = []
        self.param_needs_grad = param_needs_grad  # {id(param): bool}
        self.input_needs_grad = input_needs_grad
        self.param_grads = {}

    def needs(self, p):
        return p is not None and self.param_needs_grad.get(id(p), False)

    def add_param_grad(self, p, g):
        k = id(p)
        self.param_grads[k] = g if k not in self.param_grads else self.param_grads[k] + g

    def backward(self):
        for step in reversed(self.steps):
            step()


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


# ------------------------------------------------------------------------------------------------ fused steps
def step_conv(tape: Tape, b: Buf, m, act=ACT_NONE, slope=0.0) -> Buf:
    """conv (+bias, + optional epilogue activation).  Output: raw Buf (act NONE) or activation Buf."""
    op = m.conv_op()
    x = b.t
    y = ops.conv_forward(op, x, m.weight, m.bias, act, slope)
    out = Buf(y, 0, m.out_channels, b.is_3d, raw=(act == ACT_NONE))
    weight, bias = m.weight, m.bias
    out.want_dbias = tape is not None and tape.needs(bias)

    def bwd():
        g = out.grad
        if g is None:
            return
        out.grad = None
        db = out.dbias
        out.dbias = None
        if act != ACT_NONE:
            if tape.needs(bias):
                db = ops.zeros((y.shape[-1],), y.device)
            g = ops.act_backward(g, y, act, slope, dbias=db)  # fp32 d_buf -> bf16 d_raw (+ bias gradient)
        if tape.needs(weight):
            tape.add_param_grad(weight, op.run_wgrad(x, g, weight.shape))
        if tape.needs(bias):
            tape.add_param_grad(bias, db[:op.cout] if db is not None else ops.colsum(g, op.cout))
        if b.grad is not None or b.needs_grad_flag:
            b.grad = op.run_dgrad(g, weight, x.shape, into=b.grad)

    if tape is not None:
        tape.steps.append(bwd)
    return out


def step_norm_act(tape: Tape, raw: Buf, norm: bool, act: int, slope: float, out_pad: int, eps: float,
                  residual: Optional[Buf] = None) -> Buf:
    """[instance-norm] + activation [+ residual] + reflection border of the result."""
    if not raw.raw and not (norm is False and act == ACT_NONE):
        raise NotImplementedError("normalisation of a non-convolution output")
    if not raw.raw:
        raise NotImplementedError("explicit border copy of an activation buffer (no reference network needs it)")
    t, stats = ops.norm_act_forward(raw.t, residual.t if residual is not None else None,
                                    residual.pad if residual is not None else 0, norm, act, slope, out_pad, eps)
    out = Buf(t, out_pad, raw.channels, raw.is_3d)
    x = raw.t

    def bwd():
        g = out.grad
        if g is None:
            return
        out.grad = None
        dres = None
        if residual is not None and residual.needs_grad_flag:
            if residual.grad is not None:
                raise NotImplementedError("residual gradient buffer already exists (unsupported topology)")
            dres = ops.zeros(residual.t.shape, x.device)
            residual.grad = dres
        if raw.want_dbias and raw.needs_grad_flag:
            raw.dbias = ops.zeros((x.shape[-1],), x.device)
        raw.grad = ops.norm_act_backward(x, stats, t, g, norm, act, slope, out_pad, eps, dres32=dres,
                                         res_pad=residual.pad if residual is not None else 0,
                                         need_draw=raw.needs_grad_flag, dbias=raw.dbias)

    if tape is not None:
        tape.steps.append(bwd)
    return out


def run_sequence(tape: Tape, mods: Sequence[nn.Module], b: Buf, final_pad: int = 0,
                 residual: Optional[Buf] = None) -> Buf:
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
                # the producer did not materialise the border: copy-with-border pass
                b = step_norm_act(tape, b, False, ACT_NONE, 0.0, m.pad_amount, 1e-5)
            i += 1
            pending_pad = m.pad_amount
            if i >= n or not isinstance(mods[i], _ConvMixin):
                raise RuntimeError("ReflectionPad must be followed by a convolution")
            continue
        if isinstance(m, _ConvMixin):
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
            nxt = first_pad(mods[j:]) if j < n else final_pad  # border wanted by the consumer of this group
            res = residual if (j >= n and residual is not None) else None
            if norm or res is not None or nxt > 0:
                raw = step_conv(tape, b, m)
                b = step_norm_act(tape, raw, norm, act_id, slope, nxt, mods[i + 1].eps if norm else 1e-5, res)
            else:
                b = step_conv(tape, b, m, act_id, slope)  # bias + activation in the convolution epilogue
            i = j
            continue
        if hasattr(m, "gb_run"):
            nxt = first_pad(mods[i + 1:]) if i + 1 < n else final_pad
            b = m.gb_run(tape, b, nxt)
            i += 1
            continue
        if isinstance(m, nn.Identity):
            i += 1
            continue
        raise NotImplementedError(f"run_sequence: unsupported layer {type(m).__name__} at position {i}")
    return b


class NetworkFn(torch.autograd.Function):
    """One network = one autograd node. forward(x NC(D)HW fp32) -> NC(D)HW fp32."""

    @staticmethod
    def forward(ctx, mods, x, *params):
        needs = {id(p): ctx.needs_input_grad[2 + k] for k, p in enumerate(params)}
        record = any(ctx.needs_input_grad[1:])
        tape = Tape(needs, ctx.needs_input_grad[1]) if record else None
        act = ACT_NONE
        if len(mods) and isinstance(mods[-1], Tanh):
            mods, act = mods[:-1], ACT_TANH
        pad = first_pad(mods)
        ctx.arena_key = (id(mods[0]) if len(mods) else 0, tuple(x.shape), tuple(ctx.needs_input_grad))
        ops.arena_begin(("fwd",) + ctx.arena_key, x.device)
        b0 = Buf(ops.to_channels_last(x, pad), pad, x.shape[1], x.dim() == 5)
        b0.needs_grad_flag = bool(ctx.needs_input_grad[1])
        b = run_sequence(tape, mods, b0)
        if b.pad != 0:
            raise RuntimeError("cannot export a bordered buffer")
        y = ops.from_channels_last(b.t, b.channels, b.is_3d, act)
        ops.arena_end(("fwd",) + ctx.arena_key)
        ctx.tape, ctx.params, ctx.b0, ctx.b_last, ctx.act = tape, params, b0, b, act
        ctx.in_shape = tuple(x.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        tape, b0, b = ctx.tape, ctx.b0, ctx.b_last
        tape.param_grads = {}
        ops.arena_begin(("bwd",) + ctx.arena_key, dy.device)
        b.grad = ops.from_channels_last_backward(dy, b.t.shape, b.channels, pre=b.t if ctx.act == ACT_TANH else None,
                                                 fp32=not b.raw)
        tape.backward()
        dx = None
        if ctx.needs_input_grad[1] and b0.grad is not None:
            dx = ops.to_channels_last_backward(b0.grad, b0.pad, ctx.in_shape)
        b0.grad = None
        ops.arena_end(("bwd",) + ctx.arena_key)
        grads = [tape.param_grads.get(id(p)) for p in ctx.params]
        tape.param_grads = {}
        return (None, dx, *grads)


def run_network(net: nn.Module, mods: Sequence[nn.Module], x: torch.Tensor) -> torch.Tensor:
    """NC(D)HW fp32 in -> NC(D)HW fp32 out through the fused kernels (a trailing nn.Tanh is evaluated in fp32
    while the result is exported)."""
    params = [p for p in net.parameters()]
    return NetworkFn.apply(flatten_modules(mods), x, *params)
