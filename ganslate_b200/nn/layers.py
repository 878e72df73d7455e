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
from torch.nn.modules.utils import _triple

import sys

from .. import ops
from .._cabi import ACT_LEAKY, ACT_NONE, ACT_PRELU, ACT_RELU, ACT_TANH
from . import fp32_mode


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
            k3 = _t3(self.kernel_size)
            if k3[0] * k3[1] * k3[2] > ops._cabi.GB_MAX_TAPS and not self._transposed:
                # Resnet3D's 7x7x7 layers: kd depth slabs of kh*kw taps each (ops.SlabConv)
                op = ops.SlabConv(self.in_channels, self.out_channels, k3, _t3(self.stride), _p3(self.padding))
            else:
                op = ops.ConvOp(self.in_channels, self.out_channels, k3, _t3(self.stride), _p3(self.padding),
                                transposed=self._transposed,
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


class SeparableConv3d(nn.Module):
    """ganslate/nn/separable.py:5-40: an in-plane (1, k, k) convolution followed by a through-plane (k, 1, 1) one
    (both dense over channels, both biased); attribute names = the reference's state_dict keys."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        k, s, p = _triple(kernel_size), _triple(stride), _triple(padding)
        self.conv_depthwise = Conv3d(in_channels, out_channels, kernel_size=(1, k[1], k[2]), stride=(1, s[1], s[2]),
                                     padding=(0, p[1], p[2]), bias=bias)
        self.conv_pointwise = Conv3d(out_channels, out_channels, kernel_size=(k[0], 1, 1), stride=(s[0], 1, 1),
                                     padding=(p[0], 0, 0), bias=bias)
        self.out_channels = out_channels

    def gb_pair(self):
        return self.conv_depthwise, self.conv_pointwise

    def forward(self, x):  # noqa: D401
        raise RuntimeError("ganslate_b200 layers are executed through their network's fused forward")


class SeparableConvTranspose3d(nn.Module):
    """ganslate/nn/separable.py:43-83 (transposed counterpart)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        k, s, p = _triple(kernel_size), _triple(stride), _triple(padding)
        self.conv_transp_depthwise = ConvTranspose3d(in_channels, out_channels, kernel_size=(1, k[1], k[2]),
                                                     stride=(1, s[1], s[2]), padding=(0, p[1], p[2]), bias=bias)
        self.conv_transp_pointwise = ConvTranspose3d(out_channels, out_channels, kernel_size=(k[0], 1, 1),
                                                     stride=(s[0], 1, 1), padding=(p[0], 0, 0), bias=bias)
        self.out_channels = out_channels

    def gb_pair(self):
        return self.conv_transp_depthwise, self.conv_transp_pointwise

    def forward(self, x):  # noqa: D401
        raise RuntimeError("ganslate_b200 layers are executed through their network's fused forward")


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


class ReplicationPad3d(_Marker, nn.ReplicationPad3d):
    """Materialised by a streaming copy (csrc/pad.cu): a TMA box cannot clamp its coordinates."""

    @property
    def pads_zyx(self):
        l, r, t, b, f, k = self.padding  # (left, right, top, bottom, front, back)
        if not (l == r and t == b and f == k):
            raise NotImplementedError("asymmetric ReplicationPad3d")
        return int(f), int(t), int(l)


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


class PReLU(_Marker, nn.PReLU):
    """Per-channel learnable slopes (V-Net, ganslate/nn/generators/vnet/vnet3d.py:160,196,231,251)."""


# Storages that received a gradient during the backward pass that is running: the pass clears them when it ends, so
# no fp32 gradient buffer outlives its backward and a second backward over the same tape (retain_graph=True with
# RELEASE_TAPE off) starts from clean gradients instead of accumulating onto the previous pass's.
_LIVE_GRADS = []
_WGRAD_STREAMS = {}   # compute stream -> its weight-gradient side stream (ops.WGRAD_STREAM)

# After its backward a network's tape (the closures that hold every bf16 activation of the forward pass) is dropped,
# as autograd frees its saved tensors: a second backward through the same forward then raises, like torch's
# "Trying to backward through the graph a second time".  Set to False to keep tapes (retain_graph=True users).
RELEASE_TAPE = True


class Storage:
    """One channels-last bf16 allocation (N, D, H+2*pad, W+2*pad, C) plus, during backward, its gradient.

    raw=True marks a bare convolution output: its gradient is the bf16 MMA operand of dgrad / wgrad and is produced
    in one piece by the consumer's backward.  Every other buffer is an activation: its gradient is an FP32 tensor of
    the same shape into which every consumer ACCUMULATES (convolution dgrad epilogues, residual branches, channel
    slices of concatenations), so arbitrary fan-out / channel splits need no extra add kernels."""
    __slots__ = ("t", "pad", "raw", "_grad", "consumers")

    def __init__(self, t, pad, raw=False):
        self.t, self.pad, self.raw = t, pad, raw
        self._grad = None
        self.consumers = 0  # how many tape steps read this storage (decides zero-init vs overwrite in backward)

    @property
    def grad(self):
        return self._grad

    @grad.setter
    def grad(self, g):
        if g is not None and self._grad is None:
            _LIVE_GRADS.append(self)
        self._grad = g


def _end_backward(ctx):
    """Common tail of the three Functions' backward: drop every gradient buffer of the pass and (RELEASE_TAPE) the
    tape with the activations it holds."""
    for st in _LIVE_GRADS:
        st._grad = None
    del _LIVE_GRADS[:]
    if RELEASE_TAPE:
        ctx.tape.steps = None
        ctx.tape = ctx.b0 = ctx.b_last = None
        if hasattr(ctx, "sink"):
            ctx.sink = None


def _begin_backward(ctx):
    if ctx.tape is None or ctx.tape.steps is None:
        raise RuntimeError("ganslate_b200: backward through a network a second time -- its tape was released after "
                           "the first backward (set ganslate_b200.nn.layers.RELEASE_TAPE = False to keep tapes "
                           "for retain_graph=True)")
    del _LIVE_GRADS[:]


class Buf:
    """Channel slice [c0, c0 + cw) of a Storage; `channels` logical channels (cw = channels rounded up to 8)."""
    __slots__ = ("st", "c0", "cw", "channels", "is_3d", "needs_grad_flag", "want_dbias", "dbias", "stats", "bias_param")

    def __init__(self, t, pad, channels, is_3d, raw=False, st=None, c0=0, cw=None):
        self.st = st if st is not None else Storage(t, pad, raw)
        self.c0 = c0
        self.cw = cw if cw is not None else self.st.t.shape[-1]
        self.channels, self.is_3d = channels, is_3d
        self.needs_grad_flag = True  # False only for a network input that does not require grad
        self.want_dbias = False      # the producing conv's bias needs a gradient (fused into the consumer's backward)
        self.dbias = None
        self.bias_param = None       # that bias (direct parameter gradients: ops.DIRECT_PARAM_GRAD)
        self.stats = None            # InstanceNorm statistics written by the producing convolution's epilogue

    # -- forward-side accessors
    @property
    def t(self):
        return self.st.t

    @property
    def pad(self):
        return self.st.pad

    @property
    def raw(self):
        return self.st.raw

    @property
    def full(self):
        return self.c0 == 0 and self.cw == self.st.t.shape[-1]

    def slice(self, c0, channels):
        """Sub-slice (channel offsets relative to this Buf)."""
        return Buf(None, None, channels, self.is_3d, st=self.st, c0=self.c0 + c0, cw=ops.pad8(channels))

    def view(self):
        """Interior view of the slice."""
        return ops.make_view(self.st.t, self.st.pad, self.c0, self.cw)

    def plain_view(self):
        """The slice of the WHOLE allocation (border included) as a plain tensor -- what a convolution reads."""
        return ops.make_view(self.st.t, 0, self.c0, self.cw)

    # -- backward-side accessors
    def has_grad(self):
        return self.st.grad is not None

    def grad_tensor(self):
        """FP32 gradient of the whole storage, zero-initialised on first use (consumers accumulate)."""
        if self.st.grad is None:
            self.st.grad = ops.zeros(self.st.t.shape, self.st.t.device)
        return self.st.grad

    def grad_view(self):
        return ops.make_view(self.grad_tensor(), self.st.pad, self.c0, self.cw)

    def grad_plain_view(self):
        return ops.make_view(self.grad_tensor(), 0, self.c0, self.cw)

    @property
    def grad(self):  # used by NetworkFn for the input / seeded gradients
        return self.st.grad

    @grad.setter
    def grad(self, g):
        self.st.grad = g


class Tape:
    """Reverse-mode tape of fused steps; every step is a closure that consumes the gradient of its output
    Buf(s) and produces / accumulates the gradients of its inputs."""

    def __init__(self, param_needs_grad, input_needs_grad):
        self.steps = []
        self.param_needs_grad = param_needs_grad  # {id(param): bool}
        self.input_needs_grad = input_needs_grad
        self.param_grads = {}
        self.unpack = ops.UnpackQueue()  # weight-gradient workspaces -> PyTorch layout, one launch per pass
        self.direct = ops.DIRECT_PARAM_GRAD  # gradients go straight into param.grad (ops.DIRECT_PARAM_GRAD)
        self._queued = set()             # direct mode: parameters whose first gradient waits in the unpack queue

    def needs(self, p):
        return p is not None and self.param_needs_grad.get(id(p), False)

    def wgrad_stream(self, device):
        """Opt-in (GB_WGRAD_STREAM=1): weight gradients run on a side stream of the stream the backward pass is on.
        Nothing in the backward chain waits for a weight gradient -- only the parameter gradients handed back at the
        end do -- so its one-wave kernel overlaps the data-gradient / InstanceNorm kernels of the layers that follow.
        The pass joins the side stream before it flushes the weight-gradient copies (Tape.backward)."""
        if not ops.WGRAD_STREAM or torch.device(device).type != "cuda":
            return None
        cur = torch.cuda.current_stream()
        side = _WGRAD_STREAMS.get(cur.cuda_stream)
        if side is None:
            side = _WGRAD_STREAMS[cur.cuda_stream] = torch.cuda.Stream(device)
        self._side = (cur, side)
        return side

    def grad_target(self, p, numel=None):
        """Direct mode: the existing `p.grad` a kernel may accumulate into (None: produce a fresh gradient)."""
        if not self.direct or p is None:
            return None
        g = p.grad
        if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != p.shape:
            return None
        if numel is not None and g.numel() != numel:
            return None
        if id(p) in self._queued:  # its first contribution has not left the queue yet: order the two launches
            self.unpack.flush()
            self._queued.clear()
        return g

    def add_param_grad(self, p, g):
        if self.direct:
            if p.grad is None:
                g = g.detach()
                p.grad = g if (g.shape == p.shape and g.is_contiguous()) else g.reshape(p.shape).contiguous()
                self._queued.add(id(p))
            elif g.data_ptr() != p.grad.data_ptr():
                self.unpack.flush()
                self._queued.clear()
                p.grad.add_(g.reshape(p.shape))
            return  # (same storage: a kernel accumulated into p.grad already)
        k = id(p)
        if k in self.param_grads:
            self.unpack.flush()  # the earlier contribution may still be waiting in the queue
            g = self.param_grads[k] + g
        self.param_grads[k] = g

    def backward(self):
        if RELEASE_TAPE:
            # pop as we go: a step's closure (and with it the activations only it still references) dies as soon as
            # its gradient has been propagated, as autograd frees saved tensors node by node
            steps, self.steps = self.steps, None
            while steps:
                steps.pop()()
        else:
            for step in reversed(self.steps):
                step()
        side = self.__dict__.pop("_side", None)
        if side is not None:
            side[0].wait_stream(side[1])
        self.unpack.flush()
        self._queued.clear()


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
def step_conv(tape: Tape, b: Buf, m, act=ACT_NONE, slope=0.0, want_stats=False, as_activation=False) -> Buf:
    """conv (+bias, + optional epilogue activation) reading the whole (bordered) allocation of `b`'s channel slice.
    Output: raw Buf (act NONE) or activation Buf.  want_stats: an InstanceNorm follows -- its statistics are
    accumulated by the convolution epilogue (Buf.stats) instead of a separate pass over the output.
    as_activation: the output feeds another convolution directly (SeparableConv3d: depthwise -> pointwise), so it
    is an activation buffer whose FP32 gradient the consumer accumulates and this step converts to the bf16 operand."""
    if ops.FP32_MODE:
        return fp32_mode.step_conv(sys.modules[__name__], tape, b, m, act, slope, want_stats, as_activation)
    op = m.conv_op()
    dev = b.t.device
    ops._require_cuda(b.t, "convolution input")
    stats = ops.zeros((b.t.shape[0], op.cout_pad, 2), dev) if (want_stats and act == ACT_NONE) else None
    xin = None
    if getattr(op, "widen_input", False):
        # ops.WIDEN_INPUT: the layer reads a copy of its 8-channel operand with 8 zero channels appended (kept for the
        # weight gradient); one streaming pass over a 1-4 channel network input
        xin = torch.nn.functional.pad(b.st.t[..., b.c0:b.c0 + b.cw], (0, op.cin_pad - b.cw))
    in_view = (lambda: ops.make_view(xin)) if xin is not None else b.plain_view
    y = op.run_fwd(in_view(), dev, m.weight, m.bias, act, slope, stats=stats)
    out = Buf(y, 0, m.out_channels, b.is_3d, raw=(act == ACT_NONE and not as_activation))
    out.stats = stats
    weight, bias = m.weight, m.bias
    out.want_dbias = tape is not None and tape.needs(bias)
    out.bias_param = bias  # direct mode: the InstanceNorm backward adds its bias-gradient sums to bias.grad
    b.st.consumers += 1

    def bwd():
        g = out.st.grad
        if g is None:
            return
        out.st.grad = None
        db = out.dbias
        out.dbias = None
        if act != ACT_NONE or as_activation:
            if tape.needs(bias):
                db = tape.grad_target(bias, y.shape[-1])
                if db is None:
                    db = ops.zeros((y.shape[-1],), dev)
            g = ops.act_backward(ops.make_view(g), y, act, slope, dbias=db)  # fp32 d_buf -> bf16 d_raw (+ bias grad)
        gv = ops.make_view(g)
        if op.bwd_window and (tape.needs(weight) or b.needs_grad_flag):
            # gradient-side pixel windows (ops.ConvOp.bwd_window): dOut rows with zero pixels on both sides
            gv = torch.nn.functional.pad(g, (0, 0, ops.BWD_BORDER, ops.BWD_BORDER))
        if tape.needs(weight):
            tgt = tape.grad_target(weight)
            side = tape.wgrad_stream(dev) if tgt is None else None
            if side is not None:
                cur = torch.cuda.current_stream()
                side.wait_stream(cur)           # the gradient `g` (and everything before it) is ready
                with torch.cuda.stream(side):
                    dw = op.run_wgrad(in_view(), gv, weight.shape, dev, pending=tape.unpack)
                for t in (g, b.st.t, xin):      # read on the side stream: their memory must not be reused before it is done
                    if torch.is_tensor(t):
                        t.record_stream(side)
                tape.add_param_grad(weight, dw)
            else:
                tape.add_param_grad(weight, op.run_wgrad(in_view(), gv, weight.shape, dev, pending=tape.unpack,
                                                         dw_out=tgt, accumulate=tgt is not None))
        if tape.needs(bias):
            tape.add_param_grad(bias, db[:op.cout] if db is not None else ops.colsum(g, op.cout))
        if b.needs_grad_flag:
            if not b.has_grad() and b.st.consumers == 1 and b.full and not getattr(op, "accumulates_dgrad", False):
                b.st.grad = torch.empty(b.st.t.shape, dtype=torch.float32, device=dev)  # sole consumer: overwrite
                op.run_dgrad(gv, weight, b.grad_plain_view(), accumulate=False)
            else:
                op.run_dgrad(gv, weight, b.grad_plain_view(), accumulate=True)

    if tape is not None:
        tape.steps.append(bwd)
    return out


def step_conv_any(tape: Tape, b: Buf, m, act=ACT_NONE, slope=0.0, want_stats=False) -> Buf:
    """step_conv for a plain or a separable (two chained convolutions) layer."""
    if hasattr(m, "gb_pair"):
        first, second = m.gb_pair()
        b = step_conv(tape, b, first, as_activation=True)
        return step_conv(tape, b, second, act, slope, want_stats)
    return step_conv(tape, b, m, act, slope, want_stats)


def step_norm_act(tape: Tape, x: Buf, norm: bool, act: int, slope: float, out_pad: int, eps: float,
                  residual: Optional[Buf] = None, prelu=None, res_before_act=False, out_scale=1.0,
                  out: Optional[Buf] = None) -> Buf:
    """[instance-norm] + activation [+ residual] + reflection border of the result.

    x: raw convolution output, or (norm=False) any activation Buf; out: optional existing Buf (channel slice of a
    concatenation buffer) to write into instead of a fresh allocation; prelu: nn.PReLU module for learnable slopes."""
    if ops.FP32_MODE:
        return fp32_mode.step_norm_act(sys.modules[__name__], tape, x, norm, act, slope, out_pad, eps, residual, prelu,
                                       res_before_act, out_scale, out)
    dev = x.t.device
    ops._require_cuda(x.t, "normalisation input")
    if norm and not x.raw and (residual is not None and res_before_act):
        raise NotImplementedError("normalisation of an activation buffer with a residual before the activation")
    if out is None:
        N, D, H, W, _ = x.t.shape
        H, W = H - 2 * x.pad, W - 2 * x.pad
        t = torch.empty((N, D, H + 2 * out_pad, W + 2 * out_pad, x.cw), dtype=torch.bfloat16, device=dev)
        out = Buf(t, out_pad, x.channels, x.is_3d)
    elif out.pad != out_pad:
        raise RuntimeError("destination buffer has a different reflection border")
    slopes = prelu.weight if prelu is not None else None
    if slopes is not None and slopes.numel() != x.cw:
        if slopes.numel() != x.channels:
            raise NotImplementedError("PReLU with a single shared slope")
        slopes_p = torch.zeros(x.cw, dtype=torch.float32, device=dev)
        slopes_p[:x.channels] = slopes.detach()
    else:
        slopes_p = slopes.detach() if slopes is not None else None
    stats = ops.norm_act_forward(x.view(), out.view(), residual.view() if residual is not None else None, norm, act,
                                 slope, eps, dev, prelu=slopes_p, res_before_act=res_before_act, out_scale=out_scale,
                                 stats=x.stats if (norm and x.full) else None)
    x.st.consumers += 1
    if residual is not None:
        residual.st.consumers += 1

    def bwd():
        if not out.has_grad():
            return
        gview = out.grad_view()
        need_dx = x.needs_grad_flag
        dresv = None
        if residual is not None and residual.needs_grad_flag:
            dresv = residual.grad_view()  # zero-initialised, shared with the residual's other consumers
        dprelu = ops.zeros((x.cw,), dev) if (slopes is not None and tape.needs(slopes)) else None
        if x.raw:
            if x.want_dbias and need_dx:
                x.dbias = tape.grad_target(x.bias_param, x.cw)
                if x.dbias is None:
                    x.dbias = ops.zeros((x.cw,), dev)
            seeded = x.st.grad  # gradient that arrived through a feature tap on the raw convolution output
            draw = torch.empty_like(x.t)
            ops.norm_act_backward(x.view(), stats, gview, ops.make_view(draw), norm, act, slope, eps, dev,
                                  yv=out.view() if (not norm and act != ACT_NONE) else None,
                                  resv=residual.view() if (residual is not None and res_before_act) else None,
                                  dresv=dresv, prelu=slopes_p, dprelu=dprelu, dbias=x.dbias,
                                  res_before_act=res_before_act, dres_acc=True, out_scale=out_scale, need_dx=need_dx)
            if need_dx:
                if seeded is not None:
                    draw = draw + seeded
                    if x.dbias is not None:
                        bp = x.bias_param
                        if bp is not None and bp.grad is not None and x.dbias.data_ptr() == bp.grad.data_ptr():
                            x.dbias.add_(seeded.float().sum(dim=(0, 1, 2, 3)))  # direct mode: x.dbias IS bias.grad
                        else:
                            x.dbias = x.dbias + seeded.float().sum(dim=(0, 1, 2, 3))
                x.st.grad = draw
        else:
            # activation input (copy / add / activation of existing buffers, or an InstanceNorm that is not fed by a
            # convolution -- piresnet3d.py:116): fp32 gradient, accumulated
            ops.norm_act_backward(x.view(), stats if norm else None, gview, x.grad_view(), norm, act, slope, eps, dev,
                                  resv=residual.view() if (residual is not None and res_before_act) else None,
                                  dresv=dresv, prelu=slopes_p, dprelu=dprelu, res_before_act=res_before_act,
                                  dres_acc=True, dx_fp32_acc=True, out_scale=out_scale, need_dx=need_dx)
        if dprelu is not None:
            tape.add_param_grad(slopes, dprelu[:slopes.numel()].reshape(slopes.shape))
        if out.full:
            out.st.grad = None   # (a channel slice of a concatenation buffer shares its gradient with other producers)

    if tape is not None:
        tape.steps.append(bwd)
    return out


def step_replicate_pad(tape: Tape, b: Buf, pads) -> Buf:
    """ReplicationPad3d(b) as a new plain buffer (N, D+2pz, H+2py, W+2px, C); backward folds the FP32 gradient of the
    padded buffer onto `b`'s gradient (accumulated, like every other consumer of an activation buffer)."""
    if ops.FP32_MODE:
        return fp32_mode.step_replicate_pad(sys.modules[__name__], tape, b, pads)
    pz, py, px = pads
    if b.raw:
        raise NotImplementedError("replicate padding of a raw convolution output")
    dev = b.t.device
    N, D, Hb, Wb, _ = b.t.shape
    H, W = Hb - 2 * b.pad, Wb - 2 * b.pad
    t = torch.empty((N, D + 2 * pz, H + 2 * py, W + 2 * px, b.cw), dtype=torch.bfloat16, device=dev)
    out = Buf(t, 0, b.channels, b.is_3d)
    ops.replicate_pad_forward(b.view(), out.view(), pads)
    b.st.consumers += 1

    def bwd():
        if not out.has_grad():
            return
        g = out.st.grad
        out.st.grad = None
        if b.needs_grad_flag:
            ops.replicate_pad_backward(ops.make_view(g), b.grad_view(), pads)

    if tape is not None:
        tape.steps.append(bwd)
    return out


def run_sequence(tape: Tape, mods: Sequence[nn.Module], b: Buf, final_pad: int = 0,
                 residual: Optional[Buf] = None, taps=None, sink=None) -> Buf:
    """Execute a list of layers on buffer `b`, fusing pad/conv/norm/activation groups.

    final_pad: reflection border wanted on the last produced buffer (what follows this sequence).
    residual:  added to the output of the LAST norm group (ResidualBlock: x + conv_block(x)).
    """
    if taps is None:
        mods = flatten_modules(mods)
        taps = ()
    # taps: module indices whose output is recorded into `sink` as (index, Buf, whole_buffer) -- CUT's feature
    # extraction over `network.encoder` (ganslate/nn/gans/unpaired/cut.py:297-312)
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
            if i in taps:
                sink.append((i, b, True))  # the padded tensor itself is the feature
            i += 1
            pending_pad = m.pad_amount
            if i >= n or not isinstance(mods[i], _ConvMixin):
                raise RuntimeError("ReflectionPad must be followed by a convolution")
            continue
        if isinstance(m, ReplicationPad3d):
            b = step_replicate_pad(tape, b, m.pads_zyx)
            if i in taps:
                sink.append((i, b, True))
            i += 1
            continue
        if _is_norm(m):
            # InstanceNorm that is not fed by a convolution (first layer of Piresnet3D's coupling branch,
            # piresnet3d.py:116): statistics pass + normalisation of the activation buffer
            if m.affine or m.track_running_stats:
                raise NotImplementedError("InstanceNorm with affine/running stats (the reference uses neither)")
            j = i + 1
            act = _act_of(mods[j]) if j < n else None
            if act is not None:
                j += 1
            act_id, slope = act if act is not None else (ACT_NONE, 0.0)
            nxt = first_pad(mods[j:]) if j < n else final_pad
            res = residual if (j >= n and residual is not None) else None
            b = step_norm_act(tape, b, True, act_id, slope, nxt, m.eps, res)
            for idx in range(i, j):
                if idx in taps:
                    sink.append((idx, b, False))
            i = j
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
            tap_raw = i in taps and (norm or act is not None)
            if norm or res is not None or nxt > 0 or tap_raw:
                raw = step_conv(tape, b, m, want_stats=norm)
                if i in taps:
                    sink.append((i, raw, False))
                b = step_norm_act(tape, raw, norm, act_id, slope, nxt, mods[i + 1].eps if norm else 1e-5, res)
            else:
                b = step_conv(tape, b, m, act_id, slope)  # bias + activation in the convolution epilogue
                if i in taps:
                    sink.append((i, b, False))
            for idx in range(i + 1, j):
                if idx in taps:
                    # a tap on the norm layer sees the in-place activation that follows it (reference quirk,
                    # SURVEY.md appendix A.6); a tap on the activation sees the same tensor
                    if _is_norm(mods[idx]) and act is not None and not getattr(mods[idx + 1], "inplace", False):
                        raise NotImplementedError("feature tap on a norm layer followed by a non-inplace activation")
                    sink.append((idx, b, False))
            i = j
            continue
        if hasattr(m, "gb_run"):
            nxt = first_pad(mods[i + 1:]) if i + 1 < n else final_pad
            b = m.gb_run(tape, b, nxt)
            if i in taps:
                sink.append((i, b, False))
            i += 1
            continue
        if isinstance(m, nn.Identity):
            i += 1
            continue
        raise NotImplementedError(f"run_sequence: unsupported layer {type(m).__name__} at position {i}")
    return b


# network-boundary layout conversions (fp32 validation mode: plain torch ops on fp32 buffers)
def _to_cl(x, pad):
    return fp32_mode.to_channels_last(x, pad) if ops.FP32_MODE else ops.to_channels_last(x, pad)


def _from_cl(b, act):
    if ops.FP32_MODE:
        return fp32_mode.from_channels_last(b.t, b.channels, b.is_3d, act)
    return ops.from_channels_last(b.t, b.channels, b.is_3d, act)


def _from_cl_backward(dy, b, act):
    if ops.FP32_MODE:
        return fp32_mode.from_channels_last_backward(dy, b.t, b.channels, act)
    return ops.from_channels_last_backward(dy, b.t.shape, b.channels, pre=b.t if act == ACT_TANH else None, fp32=not b.raw)


def _to_cl_backward(b0, in_shape):
    if ops.FP32_MODE:
        return fp32_mode.to_channels_last_backward(b0.st.grad, b0.pad, in_shape)
    return ops.to_channels_last_backward(b0.st.grad, b0.pad, in_shape)


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
        b0 = Buf(_to_cl(x, pad), pad, x.shape[1], x.dim() == 5)
        b0.needs_grad_flag = bool(ctx.needs_input_grad[1])
        b = run_sequence(tape, mods, b0)
        if b.pad != 0:
            raise RuntimeError("cannot export a bordered buffer")
        y = _from_cl(b, act)
        ops.arena_end(("fwd",) + ctx.arena_key)
        ctx.tape, ctx.params, ctx.b0, ctx.b_last, ctx.act = tape, params, b0, b, act
        ctx.in_shape = tuple(x.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        _begin_backward(ctx)
        tape, b0, b = ctx.tape, ctx.b0, ctx.b_last
        tape.param_grads = {}
        ops.arena_begin(("bwd",) + ctx.arena_key, dy.device)
        b.st.grad = _from_cl_backward(dy, b, ctx.act)
        tape.backward()
        dx = None
        if ctx.needs_input_grad[1] and b0.st.grad is not None:
            dx = _to_cl_backward(b0, ctx.in_shape)
        b0.st.grad = None
        ops.arena_end(("bwd",) + ctx.arena_key)
        grads = [tape.param_grads.get(id(p)) for p in ctx.params]
        tape.param_grads = {}
        _end_backward(ctx)
        return (None, dx, *grads)


class RunnerFn(torch.autograd.Function):
    """Like NetworkFn, but the network topology is an arbitrary callable `runner(tape, b0) -> (Buf, export_act)`
    built from the step_* primitives (V-Net: channel splits, concatenations, additive couplings)."""

    @staticmethod
    def forward(ctx, runner, key, x, *params):
        needs = {id(p): ctx.needs_input_grad[3 + k] for k, p in enumerate(params)}
        record = any(ctx.needs_input_grad[2:])
        tape = Tape(needs, ctx.needs_input_grad[2]) if record else None
        ctx.arena_key = (key, tuple(x.shape), tuple(ctx.needs_input_grad))
        ops.arena_begin(("run_fwd",) + ctx.arena_key, x.device)
        b0 = Buf(_to_cl(x, 0), 0, x.shape[1], x.dim() == 5)
        b0.needs_grad_flag = bool(ctx.needs_input_grad[2])
        b, act = runner(tape, b0)
        if b.pad != 0 or not b.full:
            raise RuntimeError("cannot export a bordered / sliced buffer")
        y = _from_cl(b, act)
        ops.arena_end(("run_fwd",) + ctx.arena_key)
        ctx.tape, ctx.params, ctx.b0, ctx.b_last, ctx.act = tape, params, b0, b, act
        ctx.in_shape = tuple(x.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        _begin_backward(ctx)
        tape, b0, b = ctx.tape, ctx.b0, ctx.b_last
        tape.param_grads = {}
        ops.arena_begin(("run_bwd",) + ctx.arena_key, dy.device)
        b.st.grad = _from_cl_backward(dy, b, ctx.act)
        tape.backward()
        dx = None
        if ctx.needs_input_grad[2] and b0.st.grad is not None:
            dx = _to_cl_backward(b0, ctx.in_shape)
        b0.st.grad = None
        ops.arena_end(("run_bwd",) + ctx.arena_key)
        grads = [tape.param_grads.get(id(p)) for p in ctx.params]
        tape.param_grads = {}
        _end_backward(ctx)
        return (None, None, dx, *grads)


def step_channel_repeat(tape: Tape, b: Buf, n_repeats: int) -> Buf:
    """x.repeat(1, n, 1, 1, 1) on the channel axis (V-Net InputBlock residual, vnet3d.py:164-166). The tensors are
    the network INPUT (1-4 channels), so plain ATen indexing is used; backward sums the repeats."""
    c = b.channels
    out_c = c * n_repeats
    src = b.t[..., b.c0:b.c0 + c]
    t = torch.zeros(b.t.shape[:-1] + (ops.pad8(out_c),), dtype=b.t.dtype, device=b.t.device)
    t[..., :out_c] = src.repeat(1, 1, 1, 1, n_repeats)
    out = Buf(t, 0, out_c, b.is_3d)
    b.st.consumers += 1

    def bwd():
        if not out.has_grad() or not b.needs_grad_flag:
            return
        g = out.st.grad[..., :out_c]
        g = g.reshape(g.shape[:-1] + (n_repeats, c)).sum(dim=-2)
        b.grad_tensor()[..., b.c0:b.c0 + c] += g

    if tape is not None:
        tape.steps.append(bwd)
    return out


def new_like(b: Buf, channels: int) -> Buf:
    """Fresh activation buffer with the spatial shape of `b` and `channels` channels (concatenation target)."""
    N, D, Hb, Wb, _ = b.t.shape
    t = torch.empty((N, D, Hb, Wb, ops.pad8(channels)), dtype=b.t.dtype, device=b.t.device)
    return Buf(t, b.pad, channels, b.is_3d)


def run_network(net: nn.Module, mods: Sequence[nn.Module], x: torch.Tensor) -> torch.Tensor:
    """NC(D)HW fp32 in -> NC(D)HW fp32 out through the fused kernels (a trailing nn.Tanh is evaluated in fp32
    while the result is exported)."""
    params = [p for p in net.parameters()]
    ops._require_cuda(x, "network input")
    ops.ensure_packed(net)
    return NetworkFn.apply(flatten_modules(mods), x, *params)


class EncoderFn(torch.autograd.Function):
    """Runs the first layers of a network and returns the features recorded at the module indices `taps`
    (NC(D)HW fp32), with gradients flowing back into the network through every tap."""

    @staticmethod
    def forward(ctx, mods, taps, x, *params):
        if ops.FP32_MODE:
            raise NotImplementedError("fp32 validation mode does not cover feature taps (CUT)")
        needs = {id(p): ctx.needs_input_grad[3 + k] for k, p in enumerate(params)}
        record = any(ctx.needs_input_grad[2:])
        tape = Tape(needs, ctx.needs_input_grad[2]) if record else None
        taps = tuple(sorted(taps))
        mods = list(mods)[:taps[-1] + 1]
        if any(isinstance(m, nn.Sequential) for m in mods):
            raise NotImplementedError("feature taps need a flat encoder list")
        pad = first_pad(mods)
        ctx.arena_key = (id(mods[0]), tuple(x.shape), tuple(ctx.needs_input_grad), taps)
        ops.arena_begin(("enc_fwd",) + ctx.arena_key, x.device)
        b0 = Buf(ops.to_channels_last(x, pad), pad, x.shape[1], x.dim() == 5)
        b0.needs_grad_flag = bool(ctx.needs_input_grad[2])
        sink = []
        run_sequence(tape, mods, b0, taps=set(taps), sink=sink)
        outs = []
        for idx, b, whole in sink:
            outs.append(ops.from_channels_last(b.t, b.channels, b.is_3d, ACT_NONE, 0 if whole else b.pad))
        ops.arena_end(("enc_fwd",) + ctx.arena_key)
        if len(outs) != len(taps):
            raise RuntimeError(f"feature taps {taps} produced {len(outs)} features")
        ctx.tape, ctx.params, ctx.b0, ctx.sink = tape, params, b0, sink
        ctx.in_shape = tuple(x.shape)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        _begin_backward(ctx)
        tape, b0 = ctx.tape, ctx.b0
        tape.param_grads = {}
        ops.arena_begin(("enc_bwd",) + ctx.arena_key, douts[0].device if douts[0] is not None else b0.t.device)
        for (idx, b, whole), dy in zip(ctx.sink, douts):
            if dy is None:
                continue
            g = ops.from_channels_last_backward(dy, b.t.shape, b.channels, fp32=not b.raw, pad=0 if whole else b.pad)
            b.st.grad = g if b.st.grad is None else b.st.grad + g
        tape.backward()
        dx = None
        if ctx.needs_input_grad[2] and b0.st.grad is not None:
            dx = ops.to_channels_last_backward(b0.st.grad, b0.pad, ctx.in_shape)
        b0.st.grad = None
        ops.arena_end(("enc_bwd",) + ctx.arena_key)
        grads = [tape.param_grads.get(id(p)) for p in ctx.params]
        tape.param_grads = {}
        _end_backward(ctx)
        return (None, None, dx, *grads)


def run_encoder(net: nn.Module, mods: Sequence[nn.Module], x: torch.Tensor, taps) -> list:
    """Features of `x` after the modules `taps` of the list `mods` (same semantics as iterating the layers and
    recording `feat` after index i, including the in-place activation effect)."""
    params = [p for p in net.parameters()]
    ops._require_cuda(x, "network input")
    ops.ensure_packed(net)
    return list(EncoderFn.apply(list(mods), tuple(taps), x, *params))
