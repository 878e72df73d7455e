"""fp32 validation mode (`GB_FP32=1` / `ops.FP32_MODE = True`): the north star's "tf32-off fp32" parity mode.

Purpose: ONE switch under which a whole training iteration can be compared with the fp32 reference at 1e-4, so that a
kernel bug cannot hide behind bf16 rounding noise.  It is a validation mode, not a fast path:

* every convolution still runs on the SAME tcgen05 / TMA kernels through the SAME geometry (classes, taps, packed
  weight matrices, pixel windows, slabs) -- but on SPLIT operands: an fp32 tensor v is written as hi + mid + lo with
  hi = bf16(v), mid = bf16(v - hi), lo = bf16(v - hi - mid) (24 mantissa bits), and the product x * w is accumulated in
  fp32 by the kernels' `out_fp32 / accumulate` epilogue over the six significant terms
  lo*hi, hi*lo, mid*mid, mid*hi, hi*mid, hi*hi  (the dropped ones are below 2^-24 relative).  Forward, data gradient
  and weight gradient all do this, so every launch configuration the bf16 path uses is exercised;
* activations, raw convolution outputs and all gradients are stored in fp32 (same channels-last buffers, same views);
* InstanceNorm / activations / residuals / reflection and replication borders / the tanh export are evaluated by plain
  fp32 torch ops on the device, differentiated by torch autograd inside the tape step (they are not what this mode
  validates: the fused bf16 kernels have their own per-operator and teacher-forced tests).

Only the sequence / runner networks are covered (Resnet2D, PatchGAN2D/3D, Unet2D/3D, Vnet3D, Piresnet3D); CUT's feature
taps raise.  tests/test_fp32_mode_gpu.py asserts <= 1e-4 against the fp32 oracle for a whole CycleGAN iteration."""
import torch
import torch.nn.functional as F

from .. import ops
from .._cabi import ACT_LEAKY, ACT_NONE, ACT_PRELU, ACT_RELU, ACT_TANH

# the six products kept, smallest first (fp32 accumulation): (index into x parts, index into w parts)
_TERMS = ((2, 0), (0, 2), (1, 1), (1, 0), (0, 1), (0, 0))


def split3(v: torch.Tensor):
    """fp32 -> (hi, mid, lo) bf16 tensors with hi + mid + lo == v to 24 bits."""
    v = v.float()
    hi = v.to(torch.bfloat16)
    r = v - hi.float()
    mid = r.to(torch.bfloat16)
    lo = (r - mid.float()).to(torch.bfloat16)
    return hi, mid, lo


def _split_weight(w: torch.Tensor):
    """fp32 parameter -> three fp32 tensors holding its bf16 parts (ConvOp packs them without further rounding)."""
    hi, mid, lo = split3(w.detach())
    return hi.float().contiguous(), mid.float().contiguous(), lo.float().contiguous()


def _part_ops(op):
    """One ConvOp per weight part (each caches its own packed matrices), same geometry as `op`."""
    parts = op.__dict__.get("_fp32_parts")
    if parts is None:
        if isinstance(op, ops.SlabConv):
            parts = [ops.SlabConv(op.cin, op.cout, op.kernel, (1, 1, 1), (0,) + tuple(op.slabs[0].padding[1:])) for _ in range(3)]
        else:
            parts = [ops.ConvOp(op.cin, op.cout, op.kernel, op.stride, op.padding, transposed=op.transposed,
                                output_padding=op.output_padding) for _ in range(3)]
        op.__dict__["_fp32_parts"] = parts
    return parts


def conv_forward(op, x32: torch.Tensor, weight, bias):
    """x32: fp32 buffer (N, D, Hb, Wb, cin_pad) incl. any border -> fp32 raw output (N, Do, Ho, Wo, cout_pad)."""
    if isinstance(op, ops.SlabConv):
        raise NotImplementedError("fp32 validation mode: slab-decomposed (7x7x7) convolutions")
    N, D, Hb, Wb, _ = x32.shape
    od, oh, ow = op.out_extent((D, Hb, Wb))
    y = torch.empty((N, od, oh, ow, op.cout_pad), dtype=torch.float32, device=x32.device)
    yv = ops.make_view(y)
    xs = split3(x32)
    ws = _split_weight(weight)
    parts = _part_ops(op)
    for k, (ix, iw) in enumerate(_TERMS):
        parts[iw].run_fwd_into(ops.make_view(xs[ix]), ws[iw], bias if k == len(_TERMS) - 1 else None, yv, out_fp32=True,
                               accumulate=k > 0)
    return y


def conv_backward(op, x32, g32, weight, want_dx, want_dw, dx32=None):
    """g32: fp32 gradient of the raw output.  Returns (dx32 accumulated into `dx32` or fresh, dw in PyTorch layout)."""
    dev = x32.device
    gs = split3(g32)
    ws = _split_weight(weight)
    parts = _part_ops(op)
    if parts[0].bwd_window:
        # gradient-side pixel windows (ops.ConvOp.bwd_window): dOut rows with zero pixels on both sides, passed as tensors
        gops = [F.pad(g, (0, 0, ops.BWD_BORDER, ops.BWD_BORDER)) for g in gs]
    else:
        gops = [ops.make_view(g) for g in gs]
    dw = None
    if want_dw:
        xs = split3(x32)
        for ix, ig in _TERMS:
            t = parts[0].run_wgrad(ops.make_view(xs[ix]), gops[ig], weight.shape, dev)
            dw = t if dw is None else dw + t
    if want_dx:
        if dx32 is None:
            dx32 = torch.zeros(x32.shape, dtype=torch.float32, device=dev)
        dxv = ops.make_view(dx32)
        for ig, iw in _TERMS:
            parts[iw].run_dgrad(gops[ig], ws[iw], dxv, accumulate=True)
    return dx32, dw


# ------------------------------------------------------------------------------------------------ tape steps
def _interior(buf):
    t, p = buf.st.t, buf.st.pad
    return t[:, :, p:t.shape[2] - p, p:t.shape[3] - p, buf.c0:buf.c0 + buf.cw]


def _grad_interior_and_fold(g, pad):
    """Gradient on a reflection-bordered buffer -> gradient of the interior (transpose of the border copy)."""
    if pad == 0:
        return g
    x = torch.zeros(g.shape[:2] + (g.shape[2] - 2 * pad, g.shape[3] - 2 * pad) + g.shape[4:], dtype=g.dtype,
                    device=g.device).requires_grad_(True)
    with torch.enable_grad():
        y = _reflect_border(x, pad)
    (gx,) = torch.autograd.grad(y, x, g)
    return gx


def _reflect_border(x, pad):
    """(N, D, H, W, C) -> (N, D, H + 2p, W + 2p, C) with torch's reflection padding of H and W."""
    if pad == 0:
        return x
    N, D, H, W, Cc = x.shape
    t = x.permute(0, 1, 4, 2, 3).reshape(N * D, Cc, H, W)
    t = F.pad(t, (pad, pad, pad, pad), mode="reflect")
    return t.reshape(N, D, Cc, H + 2 * pad, W + 2 * pad).permute(0, 1, 3, 4, 2)


def step_conv(layers, tape, b, m, act=ACT_NONE, slope=0.0, want_stats=False, as_activation=False):
    op = m.conv_op()
    dev = b.t.device
    ops._require_cuda(b.t, "convolution input")
    x32 = b.st.t[..., b.c0:b.c0 + b.cw]
    if not x32.is_contiguous():
        x32 = x32.contiguous()
    weight, bias = m.weight, m.bias
    raw = conv_forward(op, x32, weight, bias)
    pre = raw
    if act == ACT_RELU:
        y = torch.relu(raw)
    elif act == ACT_LEAKY:
        y = F.leaky_relu(raw, slope)
    elif act == ACT_TANH:
        y = torch.tanh(raw)
    else:
        y = raw
    out = layers.Buf(y, 0, m.out_channels, b.is_3d, raw=(act == ACT_NONE and not as_activation))
    b.st.consumers += 1

    def bwd():
        g = out.st.grad
        if g is None:
            return
        out.st.grad = None
        if act == ACT_RELU:
            g = g * (pre > 0)
        elif act == ACT_LEAKY:
            g = g * torch.where(pre > 0, torch.ones_like(pre), torch.full_like(pre, slope))
        elif act == ACT_TANH:
            g = g * (1 - y * y)
        g = g.contiguous()
        want_dw, want_dx = tape.needs(weight), b.needs_grad_flag
        dx_target = None
        if want_dx and b.full:
            dx_target = b.grad_tensor()
        dx32, dw = conv_backward(op, x32, g, weight, want_dx, want_dw, dx32=dx_target)
        if want_dx and not b.full:
            b.grad_tensor()[..., b.c0:b.c0 + b.cw] += dx32
        if want_dw:
            tape.add_param_grad(weight, dw)
        if tape.needs(bias):
            tape.add_param_grad(bias, g.sum(dim=(0, 1, 2, 3))[:op.cout])

    if tape is not None:
        tape.steps.append(bwd)
    return out


def step_norm_act(layers, tape, x, norm, act, slope, out_pad, eps, residual=None, prelu=None, res_before_act=False,
                  out_scale=1.0, out=None):
    dev = x.t.device
    if out_scale == 0.0:
        out_scale = 1.0
    xin = _interior(x).detach().clone().requires_grad_(True)
    rin = _interior(residual).detach().clone().requires_grad_(True) if residual is not None else None
    slopes = prelu.weight if prelu is not None else None
    sl = None
    if slopes is not None:
        sl = torch.zeros(x.cw, dtype=torch.float32, device=dev)
        sl[:x.channels] = slopes.detach()
        sl.requires_grad_(True)
    with torch.enable_grad():
        v = xin
        if norm:
            var, mean = torch.var_mean(v, dim=(1, 2, 3), keepdim=True, correction=0)
            v = (v - mean) * torch.rsqrt(var + eps)
        if rin is not None and res_before_act:
            v = v + rin
        if act == ACT_RELU:
            v = torch.relu(v)
        elif act == ACT_LEAKY:
            v = F.leaky_relu(v, slope)
        elif act == ACT_PRELU:
            v = torch.where(v > 0, v, v * sl)
        elif act == ACT_TANH:
            v = torch.tanh(v)
        if out_scale != 1.0:
            v = v * out_scale
        if rin is not None and not res_before_act:
            v = v + rin
        y = _reflect_border(v, out_pad)
    if out is None:
        N, D, H, W, _ = xin.shape
        t = torch.empty((N, D, H + 2 * out_pad, W + 2 * out_pad, x.cw), dtype=torch.float32, device=dev)
        out = layers.Buf(t, out_pad, x.channels, x.is_3d)
    elif out.pad != out_pad:
        raise RuntimeError("destination buffer has a different reflection border")
    out.st.t[..., out.c0:out.c0 + out.cw] = y.detach()
    x.st.consumers += 1
    if residual is not None:
        residual.st.consumers += 1

    def bwd():
        if not out.has_grad():
            return
        g = out.st.grad[..., out.c0:out.c0 + out.cw]
        inputs = [xin] + ([rin] if rin is not None else []) + ([sl] if sl is not None else [])
        grads = torch.autograd.grad(y, inputs, g, allow_unused=True)
        gx = grads[0]
        k = 1
        if x.needs_grad_flag and gx is not None:
            if x.raw:
                x.st.grad = gx if x.st.grad is None else x.st.grad + gx
            else:
                p = x.st.pad
                gt = x.grad_tensor()
                gt[:, :, p:gt.shape[2] - p, p:gt.shape[3] - p, x.c0:x.c0 + x.cw] += gx
        if rin is not None:
            gr = grads[k]
            k += 1
            if residual.needs_grad_flag and gr is not None:
                p = residual.st.pad
                gt = residual.grad_tensor()
                gt[:, :, p:gt.shape[2] - p, p:gt.shape[3] - p, residual.c0:residual.c0 + residual.cw] += gr
        if sl is not None and tape.needs(slopes) and grads[k] is not None:
            tape.add_param_grad(slopes, grads[k][:slopes.numel()].reshape(slopes.shape))

    if tape is not None:
        tape.steps.append(bwd)
    return out


def step_replicate_pad(layers, tape, b, pads):
    pz, py, px = pads
    xin = _interior(b)
    N, D, H, W, Cc = xin.shape
    t = F.pad(xin.permute(0, 4, 1, 2, 3), (px, px, py, py, pz, pz), mode="replicate").permute(0, 2, 3, 4, 1).contiguous()
    out = layers.Buf(t, 0, b.channels, b.is_3d)
    b.st.consumers += 1

    def bwd():
        if not out.has_grad():
            return
        g = out.st.grad
        out.st.grad = None
        if b.needs_grad_flag:
            z = torch.zeros((N, D, H, W, Cc), dtype=torch.float32, device=g.device).requires_grad_(True)
            with torch.enable_grad():
                yy = F.pad(z.permute(0, 4, 1, 2, 3), (px, px, py, py, pz, pz), mode="replicate").permute(0, 2, 3, 4, 1)
            (gz,) = torch.autograd.grad(yy, z, g)
            p = b.st.pad
            gt = b.grad_tensor()
            gt[:, :, p:gt.shape[2] - p, p:gt.shape[3] - p, b.c0:b.c0 + b.cw] += gz

    if tape is not None:
        tape.steps.append(bwd)
    return out


def to_channels_last(x, pad):
    """NC(D)HW fp32 -> fp32 buffer (N, D, H + 2p, W + 2p, pad8(C)) with reflection border."""
    x = x.contiguous().float()
    if x.dim() == 4:
        x = x.unsqueeze(2)
    N, Cc, D, H, W = x.shape
    t = torch.zeros((N, D, H, W, ops.pad8(Cc)), dtype=torch.float32, device=x.device)
    t[..., :Cc] = x.permute(0, 2, 3, 4, 1)
    return _reflect_border(t, pad).contiguous()


def to_channels_last_backward(dbuf32, pad, shape):
    g = _grad_interior_and_fold(dbuf32, pad)[..., :shape[1]].permute(0, 4, 1, 2, 3)
    return (g if len(shape) == 5 else g[:, :, 0]).contiguous()


def from_channels_last(t, channels, is_3d, act):
    y = t[..., :channels].permute(0, 4, 1, 2, 3)
    y = y if is_3d else y[:, :, 0]
    return (torch.tanh(y) if act == ACT_TANH else y).contiguous()


def from_channels_last_backward(dout, buf, channels, act):
    """NC(D)HW gradient -> fp32 channels-last gradient of `buf` (pre-activation when the export applied tanh)."""
    g = dout.float()
    if g.dim() == 4:
        g = g.unsqueeze(2)
    g = g.permute(0, 2, 3, 4, 1)
    out = torch.zeros(buf.shape, dtype=torch.float32, device=dout.device)
    if act == ACT_TANH:
        th = torch.tanh(buf[..., :channels])
        g = g * (1 - th * th)
    out[..., :channels] = g
    return out
