"""Host-side operators over the C ABI: channels-last bf16 buffers, convolution geometry, autograd glue.

Activations inside a network are bf16 tensors of shape (N, D, H+2p, W+2p, C) -- "buffers" -- where p is a
reflection border materialised by the kernel that produced the buffer and C is padded to a multiple of 8.
A convolution that follows ReflectionPad(p) simply reads the whole buffer with padding 0, and its data
gradient is a buffer of the same (padded) shape whose border is folded back by the consumer.
"""
import ctypes as C
import itertools
import os
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch

from . import _cabi
from ._cabi import (ACT_LEAKY, ACT_NONE, ACT_PRELU, ACT_RELU, ACT_TANH, ConvParams, InBwdParams, InFwdParams,
                    PackParams, View, WgradParams)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---- zero-initialised scratch (statistics, weight-gradient workspaces, residual gradient buffers) -------------
# Every network pass needs ~100 small zeroed fp32 buffers; a fill kernel each would be ~400 launches per training
# step. A pass opens an arena keyed by (network, phase, shapes): the first time the requests are served one by one
# and their total is recorded, afterwards ONE zero fill serves them all as slices.
class _Arena:
    __slots__ = ("flat", "off", "need", "overflow")

    def __init__(self, flat):
        self.flat, self.off, self.need, self.overflow = flat, 0, 0, False


_ARENA = None
_ARENA_SIZES = {}


def arena_begin(key, device):
    global _ARENA
    n = _ARENA_SIZES.get(key, 0)
    flat = torch.zeros(n, dtype=torch.float32, device=device) if n > 0 else None
    _ARENA = _Arena(flat)
    return _ARENA


def arena_end(key):
    global _ARENA
    a, _ARENA = _ARENA, None
    if a is not None and (a.overflow or a.need != _ARENA_SIZES.get(key, 0)):
        _ARENA_SIZES[key] = a.need


def zeros(shape, device):
    """fp32 zeros, 16-byte aligned, from the open arena when it has room."""
    n = 1
    for d in shape:
        n *= int(d)
    a = _ARENA
    if a is None:
        return torch.zeros(shape, dtype=torch.float32, device=device)
    n4 = (n + 3) // 4 * 4
    a.need += n4
    if a.flat is not None and a.off + n4 <= a.flat.numel() and a.flat.device == torch.device(device):
        t = a.flat[a.off:a.off + n].view(shape)
        a.off += n4
        return t
    a.overflow = True
    return torch.zeros(shape, dtype=torch.float32, device=device)


# Opt-in (GB_DIRECT_PARAM_GRAD=1, not yet measured on a B200): the tape writes parameter gradients straight into
# `param.grad` -- the first contribution of an optimizer step becomes `.grad`, later ones (a generator used twice in
# one backward, both discriminator passes) are ACCUMULATED BY THE KERNELS (unpack accumulate flag, bias-gradient
# atomics) instead of being handed to autograd, whose AccumulateGrad then adds them with one ATen kernel per
# parameter (145 launches per CycleGAN step, profiles/r01p_launches_b8.md).  Autograd's gradient hooks do not fire in
# this mode, so it is refused under DistributedDataParallel (BaseGAN.parallelize_networks); the CUDA-graph path's
# flat-bucket all-reduce reads `.grad` directly and is unaffected.
DIRECT_PARAM_GRAD = os.environ.get("GB_DIRECT_PARAM_GRAD", "0") == "1"

# Opt-in (GB_WGRAD_STREAM=1): weight-gradient launches of a backward pass on a side stream (nn/layers.py, Tape.wgrad_stream)
WGRAD_STREAM = os.environ.get("GB_WGRAD_STREAM", "0") == "1"

# fp32 validation mode (nn/fp32_mode.py): fp32 buffers, every convolution on the same kernels with 3-way bf16-split
# operands accumulated in fp32, norm / activation steps in fp32 torch ops.  GB_FP32=1 or set at run time (tests).
FP32_MODE = os.environ.get("GB_FP32", "0") == "1"


def act_dtype():
    """Storage dtype of activation buffers."""
    return torch.float32 if FP32_MODE else torch.bfloat16


# Optional per-launch timing (bench.py roofline): a list that receives (family, work, unit, event0, event1).
PROFILE = None

# Pixel-window formulation of small-channel unit-stride convolutions (ConvOp.window / .bwd_window); tests switch it
# off for A/B.
WINDOW_CONV = True
# The gradient-side windows (ConvOp.bwd_window: data gradient and operand-swapped weight gradient of the 7x7 -> 3 output
# layer on the TMA-fed kernels) are the default since r02t: data gradient 170 -> 115 us, weight gradient 158 -> 123 us,
# whole step 393.8 -> 405.2 img/s.  GB_BWD_WINDOW=0 switches them off.
BWD_WINDOW_CONV = os.environ.get("GB_BWD_WINDOW", "1") == "1"
# Few input channels under a true 3-D kernel with many taps (the V-Net input block: 1 -> 16 channels, 5x5x5, on the full
# volume): the layer reads a 16-channel copy of its input (zero channels appended by the caller, layers.step_conv), which
# puts forward and weight gradient on the halo kernels for 32-byte pixels (csrc/igemm_halo_narrow.cu,
# igemm_wgrad_narrow.cu) instead of one 16-byte gather per pixel and tap.  GB_WIDEN_INPUT=0 keeps the 8-channel operand.
WIDEN_INPUT = os.environ.get("GB_WIDEN_INPUT", "1") == "1"
BWD_BORDER = 8  # zero pixels left and right of every dOut row in bwd_window mode (>= kw - 1)


def _call(family, work, unit, what, fn, *args):
    """Launch through the C ABI; when profiling is on, bracket the launch with CUDA events on the launch stream."""
    if PROFILE is None:
        _cabi.check(fn(*args), what)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _cabi.check(fn(*args), what)
    e1.record()
    PROFILE.append((family, work, unit, e0, e1))


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"ganslate_b200: {what} must live on a CUDA device -- there is no CPU path "
                           "(the CPU oracle lives under oracle/ and is test infrastructure only)")


def pad8(c: int) -> int:
    return (c + 7) // 8 * 8


def make_view(t: torch.Tensor, pad: int = 0, c0: int = 0, cw: int = None) -> View:
    """View of the interior of a contiguous (N, D, Hb, Wb, C) buffer with reflection border `pad`, optionally
    restricted to the channel slice [c0, c0 + cw) (c0 and cw multiples of 8: 16-byte vectors)."""
    assert t.dim() == 5 and t.is_contiguous(), (t.shape, t.stride())
    N, D, Hb, Wb, Cc = t.shape
    cw = Cc - c0 if cw is None else cw
    assert c0 % 8 == 0 and cw % 8 == 0 and c0 + cw <= Cc, (c0, cw, Cc)
    v = View()
    esz = t.element_size()
    v.ptr = t.data_ptr() + ((pad * Wb + pad) * Cc + c0) * esz
    v.sn, v.sz, v.sy, v.sx = D * Hb * Wb * Cc, Hb * Wb * Cc, Wb * Cc, Cc
    v.N, v.D, v.H, v.W, v.C, v.pad = N, D, Hb - 2 * pad, Wb - 2 * pad, cw, pad
    return v


def null_view() -> View:
    return View()


@dataclass
class DataSpec:
    """Class/tap decomposition of one implicit GEMM (see gb_conv_params)."""
    in_mul: Tuple[int, int, int]
    out_mul: Tuple[int, int, int]
    classes: List[dict]            # each: off(3), taps [(dz,dy,dx)], tap_ids [int]
    # filled by finalize()
    kpads: List[int] = field(default_factory=list)
    w_offsets: List[int] = field(default_factory=list)
    total_elems: int = 0

    def finalize(self, chans_pad: int, rows_pad: int):
        off = 0
        self.kpads, self.w_offsets = [], []
        # every class matrix uses the same row pitch (the largest class) so that one 2-D TMA map addresses them all
        kp = max(64, max((len(c["taps"]) * chans_pad + 63) // 64 * 64 for c in self.classes))
        for c in self.classes:
            self.kpads.append(kp)
            self.w_offsets.append(off)
            off += rows_pad * kp
        self.total_elems = off
        ntaps = sum(len(c["taps"]) for c in self.classes)
        if ntaps > _cabi.GB_MAX_TAPS:
            raise ValueError(f"convolution needs {ntaps} taps, the C ABI allows {_cabi.GB_MAX_TAPS}")
        if len(self.classes) > _cabi.GB_MAX_CLASSES:
            raise ValueError("too many parity classes")


def strided_spec(kernel, stride, padding) -> DataSpec:
    """out[o] = sum_r in[o*s + r - p] W[r]  (forward convolution / data gradient of a transposed convolution)."""
    taps, ids = [], []
    for i, r in enumerate(itertools.product(*[range(k) for k in kernel])):
        taps.append(tuple(r[d] - padding[d] for d in range(3)))
        ids.append(i)
    return DataSpec(tuple(stride), (1, 1, 1), [dict(off=(0, 0, 0), taps=taps, tap_ids=ids)])


def transposed_spec(kernel, stride, padding) -> DataSpec:
    """out[o] = sum_r in[(o + p - r)/s] W[r] where divisible (transposed convolution / data gradient of a
    convolution), decomposed by output parity so no zero-insertion is ever multiplied."""
    classes = []
    for par in itertools.product(*[range(s) for s in stride]):
        taps, ids = [], []
        for i, r in enumerate(itertools.product(*[range(k) for k in kernel])):
            ok = all((par[d] + padding[d] - r[d]) % stride[d] == 0 for d in range(3))
            if ok:
                taps.append(tuple((par[d] + padding[d] - r[d]) // stride[d] for d in range(3)))
                ids.append(i)
        classes.append(dict(off=par, taps=taps, tap_ids=ids))
    return DataSpec((1, 1, 1), tuple(stride), classes)


class ConvOp:
    """Geometry + packed-weight cache + launch helpers of one (transposed) convolution layer.

    Mirrors torch.nn.Conv{2,3}d / ConvTranspose{2,3}d semantics (weight layouts (Cout,Cin,k..) / (Cin,Cout,k..)),
    which is what the reference builds its networks from (e.g. ganslate/nn/generators/resnet/resnet2d.py:24-57).
    """

    def __init__(self, cin, cout, kernel, stride, padding, transposed=False, output_padding=(0, 0, 0), weight_taps=None,
                 allow_bwd_window=True):
        """weight_taps: taps per (row, channel) of the PARAMETER's layout when this op reads a depth slice
        `weight[:, :, dz]` of a larger kernel (SlabConv); default = this op's own tap count.
        allow_bwd_window: False for ops whose caller hands the gradient over as a view (SlabConv's depth slabs)."""
        self.cin, self.cout = cin, cout
        self.kernel, self.stride, self.padding = tuple(kernel), tuple(stride), tuple(padding)
        self.transposed, self.output_padding = transposed, tuple(output_padding)
        self.cin_pad, self.cout_pad = pad8(cin), pad8(cout)
        self.T = kernel[0] * kernel[1] * kernel[2]
        self.wT = wT = int(weight_taps) if weight_taps else self.T  # tap extent of the parameter's memory layout
        # (see WIDEN_INPUT) the gradient wrt the input keeps the 8-channel buffer: only `ncols = cin` columns are written
        self.widen_input = (WIDEN_INPUT and not FP32_MODE and not transposed and weight_taps is None and self.cin_pad == 8
                            and kernel[0] > 1 and self.T >= 27 and all(s == 1 for s in stride) and self.cout_pad in (16, 32))
        if self.widen_input:
            self.cin_pad = 16
        if transposed:
            self.fwd = transposed_spec(kernel, stride, padding)
            self.dgrad = strided_spec(kernel, stride, padding)
            # weight (Cin, Cout, T): fwd rows=cout, k-chans=cin ; dgrad rows=cin, k-chans=cout
            self.fwd_strides = (wT, cout * wT, 1)
            self.dgrad_strides = (cout * wT, wT, 1)
        else:
            self.fwd = strided_spec(kernel, stride, padding)
            self.dgrad = transposed_spec(kernel, stride, padding)
            # weight (Cout, Cin, T)
            self.fwd_strides = (cin * wT, wT, 1)
            self.dgrad_strides = (wT, cin * wT, 1)
        self.fwd_rows_pad = (cout + 15) // 16 * 16
        self.dgrad_rows_pad = (cin + 15) // 16 * 16
        # Pixel-window formulation of a unit-stride convolution over <= 8 channels with no padding in x (the
        # generators' 7x7 input layer behind its ReflectionPad): kw consecutive pixels x 8 channels are 112
        # contiguous bytes, read as ONE 64-"channel" pixel of a view whose pixel stride is 16 B.  A (dz, dy) pair
        # becomes one K block of 64 = (dx, c) (dx = 7: zero weights) and the TMA-fed kernels serve the layer,
        # instead of kd*kh*kw gathers of 16 B per output pixel.  Forward and weight gradient; the data gradient
        # keeps the tap formulation.
        self.window = (bool(WINDOW_CONV) and not transposed and all(s == 1 for s in self.stride) and self.cin_pad == 8
                       and 1 < kernel[2] <= 8 and self.padding[2] == 0 and self.cout_pad % 64 == 0
                       and kernel[0] * kernel[1] <= _cabi.GB_MAX_TAPS // 8
                       and (WINDOW_CONV == "force" or _cabi.lib().gb_tma_window_supported() == 1))
        self.win_pack_ids = None
        if self.window:
            kd, kh, kw = self.kernel
            taps = [(dz - self.padding[0], dy - self.padding[1], 0) for dz in range(kd) for dy in range(kh)]
            self.fwd = DataSpec((1, 1, 1), (1, 1, 1), [dict(off=(0, 0, 0), taps=taps, tap_ids=list(range(len(taps))))])
            # pack: pseudo-taps (dz, dy, dx in 0..7) of 8 channels each; dx >= kw -> zero
            self.win_pack_ids = [((dz * kh + dy) * kw + dx) if dx < kw else -1
                                 for dz in range(kd) for dy in range(kh) for dx in range(8)]
        self.fwd.finalize(64 if self.window else self.cin_pad, self.fwd_rows_pad)
        self.dgrad.finalize(self.cout_pad, self.dgrad_rows_pad)
        # wgrad: rows come from the "plain" tensor, columns (tap, c) from the gathered one
        self.wg_taps = [tuple(r[d] - padding[d] for d in range(3))
                        for r in itertools.product(*[range(k) for k in kernel])]
        if self.T > _cabi.GB_MAX_TAPS:
            raise ValueError("kernel too large")
        g_pad = self.cout_pad if transposed else self.cin_pad
        if self.window:
            self.wg_taps = [t for t in self.fwd.classes[0]["taps"]]
        self.wg_kpad = self.kernel[0] * self.kernel[1] * 64 if self.window else max(64, (self.T * g_pad + 63) // 64 * 64)
        self.wg_rows = cin if transposed else cout
        self.wg_rows_pad = (self.wg_rows + 127) // 128 * 128
        # Operand swap for unit-stride convolutions with few output channels (the generators' last 7x7 -> 3 layer,
        # PatchGAN's -> 1 layer): rows of the gradient matrix come from the INPUT channels and the columns are
        # (tap, output channel) gathered through the negated taps,
        #   dW'[c][t*cout_pad + r] = sum_q' x[q'][c] * dOut[q' - d_t][r],
        # which wastes pad128(cin) x pad64(T*cout_pad) MMA work instead of pad128(cout) x pad64(T*cin_pad).
        self.wg_swap = False
        if not transposed and all(s == 1 for s in self.stride) and not self.window:
            sw_kpad = max(64, (self.T * self.cout_pad + 63) // 64 * 64)
            sw_rows_pad = (cin + 127) // 128 * 128
            if sw_rows_pad * sw_kpad * 2 <= self.wg_rows_pad * self.wg_kpad:
                self.wg_swap = True
                self.wg_kpad, self.wg_rows, self.wg_rows_pad = sw_kpad, cin, sw_rows_pad
                self.wg_taps = [tuple(-v for v in t) for t in self.wg_taps]
        # Pixel windows on the GRADIENT side of a unit-stride convolution with <= 8 output channels (the generators'
        # 7x7 -> 3 output layer): both its data gradient and its (operand-swapped) weight gradient gather dOut
        # through the taps p - r, i.e. kw consecutive pixels of a dOut row per (dz, dy) -- with dOut copied into a
        # buffer that has BWD_BORDER zero pixels left and right of every row, the window never leaves the row and both
        # GEMMs run on the TMA-fed kernels (K blocks (j, co), j = window position <-> rx = kw - 1 - j).
        self.bwd_window = (allow_bwd_window and bool(WINDOW_CONV) and (BWD_WINDOW_CONV or WINDOW_CONV == "force") and self.wg_swap and not transposed and all(s == 1 for s in self.stride)
                           and self.cout_pad == 8 and 1 < kernel[2] <= 8 and self.cin_pad % 64 == 0
                           and self.padding[2] <= 1 and kernel[0] * kernel[1] <= _cabi.GB_MAX_TAPS // 8
                           and (WINDOW_CONV == "force" or _cabi.lib().gb_tma_window_supported() == 1))
        if self.bwd_window:
            kd, kh, kw = self.kernel
            pz, py, px = self.padding
            blocks = [(rz, ry) for rz in range(kd) for ry in range(kh)]
            # TMA x offset of a window: first tap column (px - kw + 1) shifted by the left border of the buffer
            self.bw_taps = [(pz - rz, py - ry, px - kw + 1 + BWD_BORDER) for rz, ry in blocks]
            self.bw_pack_ids = [((rz * kh + ry) * kw + (kw - 1 - j)) if j < kw else -1 for rz, ry in blocks for j in range(8)]
            self.bw_kpad = len(blocks) * 64
            self.dgrad = DataSpec((1, 1, 1), (1, 1, 1), [dict(off=(0, 0, 0), taps=self.bw_taps,
                                                              tap_ids=list(range(len(blocks))))])
            self.dgrad.finalize(64, self.dgrad_rows_pad)
            self.wg_taps, self.wg_kpad = list(self.bw_taps), self.bw_kpad
        self._packed = {}

    # ------------------------------------------------------------------ shapes
    def out_extent(self, ext):
        o = []
        for d in range(3):
            if self.transposed:
                o.append((ext[d] - 1) * self.stride[d] - 2 * self.padding[d] + self.kernel[d] + self.output_padding[d])
            else:
                o.append((ext[d] + 2 * self.padding[d] - self.kernel[d]) // self.stride[d] + 1)
        return tuple(o)

    def flops(self, in_ext, batch):
        """Algorithmic FLOPs of one pass: 2 * positions * Cin * Cout * taps, positions =
        output pixels for a convolution, input pixels for a transposed one; padded channels are not counted."""
        ext = in_ext if self.transposed else self.out_extent(in_ext)
        return 2.0 * batch * ext[0] * ext[1] * ext[2] * self.cin * self.cout * self.T

    # ------------------------------------------------------------------ packing
    def _pack_key(self, weight, which):
        return (which, weight.data_ptr(), weight._version, weight.device)

    def pack_params(self, weight: torch.Tensor, which: str):
        """(gb_pack_params, destination buffer) of one orientation; the destination is allocated once."""
        hit = self._packed.get(which)
        spec = self.fwd if which == "fwd" else self.dgrad
        rows, rows_pad = (self.cout, self.fwd_rows_pad) if which == "fwd" else (self.cin, self.dgrad_rows_pad)
        chans, chans_pad = (self.cin, self.cin_pad) if which == "fwd" else (self.cout, self.cout_pad)
        sn, sc, st = self.fwd_strides if which == "fwd" else self.dgrad_strides
        if hit is not None and hit[1].device == weight.device:
            dst = hit[1]
        else:
            dst = torch.empty(spec.total_elems, dtype=torch.bfloat16, device=weight.device)
            self._packed[which] = (None, dst)
        w = weight.detach()
        # (a depth slice weight[:, :, dz] of a larger kernel is strided: the pack strides carry its layout)
        assert (w.is_contiguous() or self.wT != self.T) and w.dtype == torch.float32
        p = PackParams()
        p.src, p.dst = w.data_ptr(), dst.data_ptr()
        p.sn, p.sc, p.st = sn, sc, st
        p.rows, p.rows_pad, p.chans, p.chans_pad = rows, rows_pad, chans, chans_pad
        p.nclass = len(spec.classes)
        if (which == "fwd" and self.window) or (which == "dgrad" and self.bwd_window):
            ids = self.win_pack_ids if which == "fwd" else self.bw_pack_ids
            p.ntaps[0], p.kpad[0], p.w_offset[0], p.tap_begin[0] = len(ids), spec.kpads[0], 0, 0
            for j, t in enumerate(ids):
                p.tap_id[j] = t
            return p, dst, rows_pad * spec.kpads[0]
        tb = 0
        for i, c in enumerate(spec.classes):
            p.ntaps[i], p.kpad[i], p.w_offset[i], p.tap_begin[i] = len(c["taps"]), spec.kpads[i], spec.w_offsets[i], tb
            for j, t in enumerate(c["tap_ids"]):
                p.tap_id[tb + j] = t
            tb += len(c["taps"])
        return p, dst, rows_pad * max(spec.kpads)

    def packed(self, weight: torch.Tensor, which: str) -> torch.Tensor:
        key = self._pack_key(weight, which)
        hit = self._packed.get(which)
        if hit is not None and hit[0] == key:
            return hit[1]
        p, dst, _ = self.pack_params(weight, which)
        _cabi.check(_cabi.lib().gb_pack_weights(C.byref(p), _stream()), "gb_pack_weights")
        self._packed[which] = (key, dst)
        return dst

    # ------------------------------------------------------------------ launches
    def _fill(self, spec: DataSpec, p: ConvParams):
        p.nclass = len(spec.classes)
        tb = 0
        for i, c in enumerate(spec.classes):
            k = p.cls[i]
            k.off[0], k.off[1], k.off[2] = c["off"]
            k.ntaps, k.tap_begin, k.kpad, k.w_offset = len(c["taps"]), tb, spec.kpads[i], spec.w_offsets[i]
            for j, t in enumerate(c["taps"]):
                p.taps[tb + j][0], p.taps[tb + j][1], p.taps[tb + j][2] = t
            tb += len(c["taps"])
        for d in range(3):
            p.in_mul[d], p.out_mul[d] = spec.in_mul[d], spec.out_mul[d]

    def _params(self, which):
        cache = self.__dict__.setdefault("_ptemplates", {})
        if which not in cache:
            p = ConvParams()
            self._fill(self.fwd if which == "fwd" else self.dgrad, p)
            cache[which] = p
        return cache[which]

    def run_fwd(self, xv: View, device, weight, bias, act=ACT_NONE, slope=0.0, stats=None):
        """xv: plain view (N,D,H,W,cin_pad) of the input (any border is part of it) -> new (N,D,Ho,Wo,cout_pad).
        stats: optional zeroed fp32 (N, cout_pad, 2) that receives the InstanceNorm statistics of the output
        (sum, sum of squares per image and channel) from the convolution epilogue."""
        assert xv.C == self.cin_pad and xv.pad == 0, (xv.C, self.cin_pad, xv.pad)
        od, oh, ow = self.out_extent((xv.D, xv.H, xv.W))
        y = torch.empty((xv.N, od, oh, ow, self.cout_pad), dtype=torch.bfloat16, device=device)
        self.run_fwd_into(xv, weight, bias, make_view(y), act, slope, stats)
        return y

    def run_fwd_into(self, xv: View, weight, bias, outv: View, act=ACT_NONE, slope=0.0, stats=None, out_fp32=False,
                     accumulate=False):
        """run_fwd into an existing view; out_fp32 / accumulate: FP32 destination that is added to (SlabConv sums the
        depth slabs of a large kernel this way)."""
        p = self._params("fwd")
        p.inp, p.out = (self.window_view(xv) if self.window else xv), outv
        p.in_c_valid = self.kernel[2] * 8 if self.window else 0
        wp = self.packed(weight, "fwd")
        p.wpacked = wp.data_ptr()
        p.bias = bias.data_ptr() if bias is not None else None
        p.ncols, p.npad = self.cout, self.fwd_rows_pad
        p.act, p.act_slope = act, slope
        p.out_fp32, p.accumulate = (1 if out_fp32 else 0), (1 if accumulate else 0)
        p.stats = stats.data_ptr() if stats is not None else None
        _call("conv_fwd", self.flops((xv.D, xv.H, xv.W), xv.N), "flop", "gb_conv_data(fwd)", _cabi.lib().gb_conv_data,
              C.byref(p), _stream())

    def bwd_window_view(self, gw: torch.Tensor) -> View:
        """gw: dOut copied into a (N, D, Ho, Wo + 2*BWD_BORDER, 8) buffer with zero borders -> the 64-"channel" window
        view over the WHOLE rows (origin = buffer start; the taps carry the border shift)."""
        N, D, H, Wb, Cc = gw.shape
        assert Cc == 8 and gw.is_contiguous()
        v = View()
        v.ptr = gw.data_ptr()
        v.sn, v.sz, v.sy, v.sx = D * H * Wb * 8, H * Wb * 8, Wb * 8, 8
        v.N, v.D, v.H, v.W, v.C, v.pad = N, D, H, Wb - self.kernel[2] + 1, 64, 0
        return v

    def run_dgrad(self, dyv, weight, outv: View, accumulate: bool):
        """dyv: bf16 view (N,D,Ho,Wo,cout_pad); the gradient wrt the input is written / accumulated into the FP32
        view outv (plain view of the input buffer's gradient, channel slice allowed).  In bwd_window mode dyv is the
        zero-bordered TENSOR (see bwd_window_view)."""
        if self.bwd_window:
            dyv = self.bwd_window_view(dyv)
        assert (self.bwd_window or dyv.C == self.cout_pad) and outv.pad == 0
        assert outv.C == (pad8(self.cin) if self.widen_input else self.cin_pad), (outv.C, self.cin_pad)
        p = self._params("dgrad")
        p.in_c_valid = self.kernel[2] * 8 if self.bwd_window else 0
        p.out_fp32, p.accumulate = 1, 1 if accumulate else 0
        p.stats = None
        p.inp, p.out = dyv, outv
        wp = self.packed(weight, "dgrad")
        p.wpacked = wp.data_ptr()
        p.bias = None
        p.ncols, p.npad = self.cin, self.dgrad_rows_pad
        p.act, p.act_slope = ACT_NONE, 0.0
        _call("conv_dgrad", self.flops((outv.D, outv.H, outv.W), outv.N), "flop", "gb_conv_data(dgrad)",
              _cabi.lib().gb_conv_data, C.byref(p), _stream())

    def wgrad_plan(self):
        """Host-side description of the weight-gradient GEMM (include/ganslate_b200.h, gb_wgrad_params) and of the
        copy of its fp32 workspace into the PyTorch weight layout; tests/test_host_logic.py evaluates it literally."""
        if self.wg_swap:
            # workspace rows = input channel c, columns = (tap, output channel r); weight layout (cout, cin, T)
            cols, cols_pad, dsr, dsc = self.cout, self.cout_pad, self.wT, self.cin * self.wT
        else:
            cols = self.cout if self.transposed else self.cin
            cols_pad = self.cout_pad if self.transposed else self.cin_pad
            dsr, dsc = cols * self.wT, self.wT
        return dict(plain_is_input=self.transposed or self.wg_swap, taps=self.wg_taps, mul=self.stride,
                    rows=self.wg_rows, rows_pad=self.wg_rows_pad, kpad=self.wg_kpad,
                    unpack=dict(dsr=dsr, dsc=dsc, dst_t=1, rows=self.wg_rows, chans=cols, chans_pad=cols_pad,
                                ntaps=self.T, kpad=self.wg_kpad))

    def window_view(self, xv: View) -> View:
        """The 8-channel plain view re-read as kw-pixel windows: C = 64 (kw*8 of them exist), pixel stride unchanged
        (8 elements), one window per output column."""
        assert xv.C == 8 and xv.sx == 8 and xv.pad == 0
        wv = View()
        C.memmove(C.byref(wv), C.byref(xv), C.sizeof(View))
        wv.C, wv.W = 64, xv.W - self.kernel[2] + 1
        return wv

    def run_wgrad(self, xv: View, dyv: View, weight_shape, device, pending=None, dw_out=None, accumulate=False):
        """fp32 weight gradient in the PyTorch layout of `weight_shape` (xv: plain input view, dyv: bf16 d_raw).
        dw_out: existing destination (a depth slice `dw[:, :, dz]` of a larger kernel's gradient, SlabConv; or
        `weight.grad` with accumulate=True: dw_out += gradient, DIRECT_PARAM_GRAD)."""
        assert not accumulate or dw_out is not None
        if self.window or self.bwd_window:
            return self._run_wgrad_window(xv, dyv, weight_shape, device, pending, dw_out, accumulate)
        plan = self.wgrad_plan()
        plain, gathered = (xv, dyv) if plan["plain_is_input"] else (dyv, xv)
        ws = zeros((self.wg_rows_pad, self.wg_kpad), device)
        cache = self.__dict__.setdefault("_ptemplates", {})
        p = cache.get("wgrad")
        if p is None:
            p = WgradParams()
            p.ntaps = self.T
            for j, t in enumerate(self.wg_taps):
                p.taps[j][0], p.taps[j][1], p.taps[j][2] = t
            for d in range(3):
                p.mul[d] = self.stride[d]
            p.rows, p.kpad, p.splits = self.wg_rows, self.wg_kpad, 0
            cache["wgrad"] = p
        p.plain, p.gathered, p.dw = plain, gathered, ws.data_ptr()
        _call("conv_wgrad", self.flops((xv.D, xv.H, xv.W), xv.N), "flop", "gb_conv_wgrad",
              _cabi.lib().gb_conv_wgrad, C.byref(p), _stream())
        dw = torch.empty(weight_shape, dtype=torch.float32, device=device) if dw_out is None else dw_out
        u = plan["unpack"]
        if pending is None and accumulate:
            pending = own = UnpackQueue()
        else:
            own = None
        if pending is not None:
            # the workspace -> PyTorch-layout copies of a whole backward pass go out in one launch (UnpackQueue)
            pending.add(ws, dw, u["dsr"], u["dsc"], u["dst_t"], u["rows"], u["chans"], u["chans_pad"], u["ntaps"], u["kpad"],
                        accumulate=accumulate)
            if own is not None:
                own.flush()
            return dw
        _cabi.check(_cabi.lib().gb_unpack_wgrad(ws.data_ptr(), dw.data_ptr(), u["dsr"], u["dsc"], u["dst_t"], u["rows"],
                                                u["chans"], u["chans_pad"], u["ntaps"], u["kpad"], _stream()),
                    "gb_unpack_wgrad")
        return dw


def _at(t: torch.Tensor, off: int) -> torch.Tensor:
    """One-element alias of `t`'s storage `off` elements behind t's first element (its data_ptr is what matters)."""
    return t.as_strided((1,), (1,), t.storage_offset() + off)


def _conv_op_window_wgrad(self, xv, dyv, weight_shape, device, pending, dw_out=None, accumulate=False):
    """Weight gradient of a pixel-window convolution: dW[r][(dz,dy)*64 + dx*8 + c] = sum_q dOut[q][r] * window[q + (dz,dy)]
    [dx*8 + c]; one unpack item per (dz, dy) K block copies its kw x cin columns into the PyTorch layout."""
    kd, kh, kw = self.kernel
    ws = zeros((self.wg_rows_pad, self.wg_kpad), device)
    cache = self.__dict__.setdefault("_ptemplates", {})
    p = cache.get("wgrad")
    if p is None:
        p = WgradParams()
        p.ntaps = len(self.wg_taps)
        for j, t in enumerate(self.wg_taps):
            p.taps[j][0], p.taps[j][1], p.taps[j][2] = t
        for d in range(3):
            p.mul[d] = 1
        p.rows, p.kpad, p.splits = self.wg_rows, self.wg_kpad, 0
        p.gathered_c_valid = kw * 8
        cache["wgrad"] = p
    if self.bwd_window:   # operand-swapped: rows = input channels, gathered = dOut windows (dyv: zero-bordered tensor)
        p.plain, p.gathered, p.dw = xv, self.bwd_window_view(dyv), ws.data_ptr()
    else:
        p.plain, p.gathered, p.dw = dyv, self.window_view(xv), ws.data_ptr()
    _call("conv_wgrad", self.flops((xv.D, xv.H, xv.W), xv.N), "flop", "gb_conv_wgrad", _cabi.lib().gb_conv_wgrad,
          C.byref(p), _stream())
    dw = torch.empty(weight_shape, dtype=torch.float32, device=device) if dw_out is None else dw_out
    own = UnpackQueue() if pending is None else pending
    wsf = ws.view(-1)
    for it in self.window_unpack_items():
        own.add(wsf[it["ws_off"]:], _at(dw, it["dst_off"]), it["dsr"], it["dsc"], it["dst_t"], it["rows"], it["chans"],
                it["chans_pad"], it["ntaps"], it["kpad"], accumulate=accumulate, keep=(ws, dw))
    if pending is None:
        own.flush()
    return dw


def _conv_op_window_unpack_items(self):
    """gb_unpack_wgrad items of a pixel-window weight gradient: K block b = (dz, dy) holds kw taps of 8 channels;
    PyTorch layout (cout, cin, kd*kh*kw): row stride cin*T, channel stride T, tap stride 1."""
    kd, kh, kw = self.kernel
    if self.bwd_window:
        # workspace rows = input channel, columns (block, j, co) with rx = kw - 1 - j: walk the taps backwards
        return [dict(ws_off=b * 64, dst_off=b * kw + kw - 1, dsr=self.wT, dsc=self.cin * self.wT, dst_t=-1,
                     rows=self.wg_rows, chans=self.cout, chans_pad=8, ntaps=kw, kpad=self.wg_kpad)
                for b in range(kd * kh)]
    return [dict(ws_off=b * 64, dst_off=b * kw, dsr=self.cin * self.wT, dsc=self.wT, dst_t=1, rows=self.wg_rows,
                 chans=self.cin, chans_pad=8, ntaps=kw, kpad=self.wg_kpad) for b in range(kd * kh)]


ConvOp._run_wgrad_window = _conv_op_window_wgrad
ConvOp.window_unpack_items = _conv_op_window_unpack_items


def _shift_z(v: View, dz: int, depth: int, elem_bytes: int) -> View:
    """The view `dz` planes deeper, `depth` planes long."""
    w = View()
    C.memmove(C.byref(w), C.byref(v), C.sizeof(View))
    w.ptr = v.ptr + dz * v.sz * elem_bytes
    w.D = depth
    return w


class SlabConv:
    """A unit-stride, depth-unpadded 3-D convolution whose kd*kh*kw taps exceed GB_MAX_TAPS (the 7x7x7 layers of
    ganslate/nn/generators/resnet/resnet3d.py:25,64: 343 taps against 128), evaluated as kd convolutions with kernel
    (1, kh, kw) over depth-shifted views of the input:  y = sum_dz conv_(1,kh,kw)(x[z + dz], W[:, :, dz]).

    Forward: the slabs accumulate into an FP32 buffer (the data kernel's out_fp32 / accumulate epilogue), which is then
    rounded to the bf16 raw output once; InstanceNorm statistics come from gb_in_stats.  Data gradient: every slab
    accumulates into its depth-shifted view of the FP32 input gradient.  Weight gradient: slab dz fills dW[:, :, dz].
    Same interface as ConvOp as far as nn/layers.py::step_conv uses it."""
    accumulates_dgrad = True   # run_dgrad always adds: the caller passes a zero-initialised gradient buffer
    bwd_window = False
    window = False
    transposed = False

    def __init__(self, cin, cout, kernel, stride, padding):
        kd, kh, kw = kernel
        if tuple(stride) != (1, 1, 1) or padding[0] != 0:
            raise NotImplementedError("convolutions with more than GB_MAX_TAPS taps need unit stride and no depth padding "
                                      "(pad with an explicit ReplicationPad3d / ReflectionPad module, as resnet3d.py does)")
        if kh * kw > _cabi.GB_MAX_TAPS:
            raise ValueError("kernel too large")
        self.cin, self.cout, self.kernel = cin, cout, tuple(kernel)
        self.cin_pad, self.cout_pad = pad8(cin), pad8(cout)
        self.T = kd * kh * kw
        self.slabs = [ConvOp(cin, cout, (1, kh, kw), (1, 1, 1), (0, padding[1], padding[2]), weight_taps=self.T,
                             allow_bwd_window=False)
                      for _ in range(kd)]

    def out_extent(self, ext):
        _, oh, ow = self.slabs[0].out_extent((1, ext[1], ext[2]))
        return (ext[0] - self.kernel[0] + 1, oh, ow)

    def flops(self, in_ext, batch):
        ext = self.out_extent(in_ext)
        return 2.0 * batch * ext[0] * ext[1] * ext[2] * self.cin * self.cout * self.T

    def run_fwd(self, xv: View, device, weight, bias, act=ACT_NONE, slope=0.0, stats=None):
        if act != ACT_NONE:
            raise NotImplementedError("epilogue activation on a slab-decomposed convolution")
        assert xv.C == self.cin_pad and xv.pad == 0
        od, oh, ow = self.out_extent((xv.D, xv.H, xv.W))
        y32 = torch.empty((xv.N, od, oh, ow, self.cout_pad), dtype=torch.float32, device=device)
        yv = make_view(y32)
        for dz, slab in enumerate(self.slabs):
            slab.run_fwd_into(_shift_z(xv, dz, od, 2), weight[:, :, dz], bias if dz == 0 else None, yv, out_fp32=True,
                              accumulate=dz > 0)
        # FP32 -> bf16 raw output (gb_in_bwd without norm / activation is a plain rounding copy)
        y = act_backward(yv, torch.empty(y32.shape, dtype=torch.bfloat16, device=device), ACT_NONE, 0.0)
        if stats is not None:
            v = make_view(y)
            _cabi.check(_cabi.lib().gb_in_stats(C.byref(v), stats.data_ptr(), _stream()), "gb_in_stats")
        return y

    def run_dgrad(self, dyv: View, weight, outv: View, accumulate: bool):
        assert accumulate, "SlabConv.run_dgrad adds to a zero-initialised gradient buffer (accumulates_dgrad)"
        for dz, slab in enumerate(self.slabs):
            slab.run_dgrad(dyv, weight[:, :, dz], _shift_z(outv, dz, dyv.D, 4), accumulate=True)

    def run_wgrad(self, xv: View, dyv: View, weight_shape, device, pending=None, dw_out=None, accumulate=False):
        dw = torch.empty(weight_shape, dtype=torch.float32, device=device) if dw_out is None else dw_out
        for dz, slab in enumerate(self.slabs):
            slab.run_wgrad(_shift_z(xv, dz, dyv.D, 2), dyv, weight_shape, device, pending, dw_out=dw[:, :, dz],
                           accumulate=accumulate)
        return dw


class UnpackQueue:
    """Weight-gradient workspaces waiting for their copy into the PyTorch layout; flush() issues ONE launch per
    GB_UNPACK_BATCH entries (the batch travels by value as a kernel parameter: CUDA-graph safe)."""

    def __init__(self):
        self.batch = _cabi.UnpackBatch()
        self.keep = []

    def add(self, ws, dw, dsr, dsc, dst_t, rows, chans, chans_pad, ntaps, kpad, accumulate=False, keep=None):
        if self.batch.count == _cabi.GB_UNPACK_BATCH:
            self.flush()
        it = self.batch.item[self.batch.count]
        it.dw, it.dst, it.dsr, it.dsc, it.dst_t = ws.data_ptr(), dw.data_ptr(), dsr, dsc, dst_t
        it.rows, it.chans, it.chans_pad, it.ntaps, it.kpad, it.accumulate = (rows, chans, chans_pad, ntaps, kpad,
                                                                             1 if accumulate else 0)
        self.batch.count += 1
        self.keep.append((ws, dw) if keep is None else keep)

    def flush(self):
        if self.batch.count:
            _call("unpack", 0, "byte", "gb_unpack_wgrad_multi", _cabi.lib().gb_unpack_wgrad_multi, C.byref(self.batch),
                  _stream())
            self.batch.count = 0
            self.keep = []


class PackGroup:
    """Every packed (bf16, class-matrix) weight of one network, refreshed by ONE launch when any parameter changed.
    The table of gb_pack_params lives in device memory; it is rebuilt only when a parameter moved."""

    def __init__(self, convs):
        self.convs = list(convs)  # [(ConvOp, weight Parameter)]
        self.table = None
        self.ptrs = None

    def _build(self):
        entries, self.items, mx = [], [], 0
        for op, w in self.convs:
            for which in ("fwd", "dgrad"):
                p, dst, n = op.pack_params(w, which)
                entries.append(bytes(p))
                self.items.append((op, w, which, dst))
                mx = max(mx, n)
        dev = self.convs[0][1].device
        self.table = torch.frombuffer(bytearray(b"".join(entries)), dtype=torch.uint8).to(dev)
        self.max_elems = mx
        self.ptrs = [w.data_ptr() for _, w in self.convs]

    def refresh(self):
        if not self.convs:
            return
        if self.ptrs is None or self.ptrs != [w.data_ptr() for _, w in self.convs]:
            self._build()
        stale = False
        for op, w, which, dst in self.items:
            hit = op._packed.get(which)
            if hit is None or hit[0] != op._pack_key(w, which) or hit[1] is not dst:
                stale = True
                break
        if not stale:
            return
        _call("pack", 0, "byte", "gb_pack_weights_multi", _cabi.lib().gb_pack_weights_multi, self.table.data_ptr(),
              len(self.items), self.max_elems, _stream())
        for op, w, which, dst in self.items:
            op._packed[which] = (op._pack_key(w, which), dst)


def ensure_packed(net):
    """Refresh the packed weights of every convolution of `net` (one launch, only when a parameter changed)."""
    g = net.__dict__.get("_gb_pack_group")
    if g is None:
        convs = []
        for m in net.modules():
            if hasattr(m, "conv_op"):
                op = m.conv_op()
                if isinstance(op, SlabConv):  # one packed matrix pair per depth slab, read from weight[:, :, dz]
                    convs += [(slab, m.weight[:, :, dz]) for dz, slab in enumerate(op.slabs)]
                else:
                    convs.append((op, m.weight))
        for _, w in convs:
            _require_cuda(w, "network parameters")
        g = PackGroup(convs)
        net.__dict__["_gb_pack_group"] = g
    g.refresh()


def colsum(t: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.empty(t.shape[-1], dtype=torch.float32, device=t.device)
    v = make_view(t)
    _cabi.check(_cabi.lib().gb_colsum(C.byref(v), out.data_ptr(), _stream()), "gb_colsum")
    return out[:n]


# ------------------------------------------------------------------------------------------------------------
# Plain (non-autograd) forward / backward helpers.  Gradient dtypes: the gradient wrt a RAW convolution output
# is bf16 (it is an MMA operand of dgrad / wgrad); the gradient wrt an ACTIVATION buffer is fp32, because
# InstanceNorm-backward subtracts its mean and bf16 rounding there costs 10-30 % error on real GAN gradients.
# ------------------------------------------------------------------------------------------------------------
def act_backward(dyv: View, y: torch.Tensor, act: int, slope: float, dbias=None) -> torch.Tensor:
    """bf16 d_raw = dy * act'(.) from the forward output y (epilogue activations); dyv: fp32 gradient view.
    dbias: optional zeroed fp32 [C] that receives the per-channel sum of d_raw (bias gradient)."""
    dx = torch.empty_like(y)
    p = InBwdParams()
    p.x, p.y, p.dy_a, p.dx = make_view(y), make_view(y), dyv, make_view(dx)
    p.act, p.act_slope, p.eps = act, slope, 1e-5
    p.dbias = dbias.data_ptr() if dbias is not None else None
    _cabi.check(_cabi.lib().gb_in_bwd(C.byref(p), _stream()), "gb_in_bwd(act)")
    return dx


def norm_act_forward(xv: View, yv: View, resv, norm, act, slope, eps, device, prelu=None, res_before_act=False,
                     out_scale=1.0, stats=None):
    """yv <- [out_scale *] act(instance_norm(xv) [+ res]) [+ res] incl. the reflection border of yv. Returns stats."""
    lib = _cabi.lib()
    nbytes = xv.N * xv.D * xv.H * xv.W * xv.C * 2
    if not norm:
        stats = None
    elif stats is None:
        stats = zeros((xv.N, xv.C, 2), device)
        _call("in_stats", nbytes, "byte", "gb_in_stats", lib.gb_in_stats, C.byref(xv), stats.data_ptr(), _stream())
    p = InFwdParams()
    p.x, p.y = xv, yv
    if resv is not None:
        p.res = resv
    p.stats = stats.data_ptr() if norm else None
    p.prelu = prelu.data_ptr() if prelu is not None else None
    p.eps, p.act, p.act_slope = eps, act, slope
    p.res_before_act, p.out_scale = 1 if res_before_act else 0, out_scale
    _call("in_fwd", nbytes * (3 if resv is not None else 2), "byte", "gb_in_fwd", lib.gb_in_fwd, C.byref(p), _stream())
    return stats


def norm_act_backward(xv: View, stats, dyv: View, dxv: View, norm, act, slope, eps, device, yv=None, resv=None,
                      dresv=None, prelu=None, dprelu=None, dbias=None, res_before_act=False, dres_acc=False,
                      dx_fp32_acc=False, out_scale=1.0, need_dx=True):
    """Backward of norm_act_forward. dyv: fp32 gradient of the (bordered) output; dxv: bf16 d_raw (or an fp32 view
    accumulated into, when the forward input was an activation buffer); dresv: fp32 gradient view of the residual."""
    lib = _cabi.lib()
    p = InBwdParams()
    p.x, p.dy_b, p.dx = xv, dyv, dxv
    if dresv is not None:
        p.dy_sum = dresv
    if resv is not None:
        p.res = resv
    p.eps = eps
    p.res_before_act, p.dy_sum_acc, p.dx_fp32_acc, p.out_scale = (1 if res_before_act else 0, 1 if dres_acc else 0,
                                                                   1 if dx_fp32_acc else 0, out_scale)
    if need_dx:
        if norm:
            bstats = zeros((xv.N * xv.C * 2 + 4,), device)  # + grid-barrier words of the single-launch path
            p.stats, p.bstats = stats.data_ptr(), bstats.data_ptr()
        elif yv is not None:
            p.y = yv
        p.act, p.act_slope = act, slope
        p.prelu = prelu.data_ptr() if prelu is not None else None
        p.dprelu = dprelu.data_ptr() if dprelu is not None else None
        p.dbias = dbias.data_ptr() if dbias is not None else None
    else:
        p.act = ACT_NONE  # only the residual branch needs the (folded) gradient
    n = xv.N * xv.D * xv.H * xv.W * xv.C
    # ALGORITHMIC bytes (read dy, read the saved x, write dx; dy is carried in fp32 here): 4 + 2 + 2 per
    # element, + 8 for the read-modify-write of a fused residual gradient.  The two-pass kernels read dy and x twice
    # (the second time mostly from L2), which is traffic, not work, and is not counted.
    nbytes = n * (8 + (8 if dresv is not None else 0))
    _call("in_bwd", nbytes, "byte", "gb_in_bwd", lib.gb_in_bwd, C.byref(p), _stream())


def replicate_pad_forward(srcv: View, dstv: View, pads):
    """dstv (plain view of a buffer with extents srcv + 2*pads) <- ReplicationPad3d(srcv); pads = (pz, py, px)."""
    nbytes = dstv.N * dstv.D * dstv.H * dstv.W * dstv.C * 2 * 2
    _call("pad", nbytes, "byte", "gb_replicate_pad_fwd", _cabi.lib().gb_replicate_pad_fwd, C.byref(srcv), C.byref(dstv),
          int(pads[0]), int(pads[1]), int(pads[2]), _stream())


def replicate_pad_backward(ddstv: View, dsrcv: View, pads):
    """dsrcv (FP32 view, accumulated) += fold of the FP32 gradient ddstv on the padded domain."""
    nbytes = ddstv.N * ddstv.D * ddstv.H * ddstv.W * ddstv.C * 4 + dsrcv.N * dsrcv.D * dsrcv.H * dsrcv.W * dsrcv.C * 8
    _call("pad", nbytes, "byte", "gb_replicate_pad_bwd", _cabi.lib().gb_replicate_pad_bwd, C.byref(ddstv), C.byref(dsrcv),
          int(pads[0]), int(pads[1]), int(pads[2]), _stream())


def to_channels_last(x: torch.Tensor, pad: int) -> torch.Tensor:
    """NC(D)HW fp32 -> bf16 buffer with reflection border `pad` (cyclegan.py:89-90 hands NCHW fp32 to the nets)."""
    _require_cuda(x, "network input")
    x = x.contiguous().float()
    if x.dim() == 4:
        N, Cc, H, W = x.shape
        D = 1
    else:
        N, Cc, D, H, W = x.shape
    out = torch.empty((N, D, H + 2 * pad, W + 2 * pad, pad8(Cc)), dtype=torch.bfloat16, device=x.device)
    v = make_view(out, pad)
    _cabi.check(_cabi.lib().gb_nchw_to_cl(x.data_ptr(), Cc, C.byref(v), None, 0, _stream()), "gb_nchw_to_cl")
    return out


def to_channels_last_backward(dbuf32: torch.Tensor, pad: int, shape) -> torch.Tensor:
    """fp32 gradient of the bordered buffer -> NC(D)HW fp32 (border folded back)."""
    dx = torch.empty(shape, dtype=torch.float32, device=dbuf32.device)
    v = make_view(dbuf32, pad, 0, pad8(shape[1]))
    _cabi.check(_cabi.lib().gb_cl_to_nchw(C.byref(v), dx.data_ptr(), shape[1], 1, ACT_NONE, 1, _stream()), "gb_cl_to_nchw")
    return dx


def from_channels_last(x: torch.Tensor, channels: int, is_3d: bool, act: int = ACT_NONE, pad: int = 0) -> torch.Tensor:
    """bf16 buffer (N,D,H+2p,W+2p,Cpad) -> NC(D)HW fp32 of its interior; act=ACT_TANH evaluates the output tanh in
    fp32."""
    N, D, Hb, Wb, Cc = x.shape
    H, W = Hb - 2 * pad, Wb - 2 * pad
    shape = (N, channels, D, H, W) if is_3d else (N, channels, H, W)
    out = torch.empty(shape, dtype=torch.float32, device=x.device)
    v = make_view(x, pad)
    _cabi.check(_cabi.lib().gb_cl_to_nchw(C.byref(v), out.data_ptr(), channels, 0, act, 0, _stream()), "gb_cl_to_nchw")
    return out


def from_channels_last_backward(dout: torch.Tensor, buf_shape, channels: int, pre=None, fp32=False,
                                pad: int = 0) -> torch.Tensor:
    """NC(D)HW fp32 gradient -> channels-last gradient (bf16 d_raw, or fp32 for an activation buffer);
    pre: the saved pre-activation when the export applied tanh; pad: the gradient is written to the interior of a
    zero-initialised bordered buffer."""
    dout = dout.contiguous().float()
    if pad > 0:
        dx = zeros(buf_shape, dout.device) if fp32 else torch.zeros(buf_shape, dtype=torch.bfloat16, device=dout.device)
    else:
        dx = torch.empty(buf_shape, dtype=torch.float32 if fp32 else torch.bfloat16, device=dout.device)
    v = make_view(dx, pad)
    pv = make_view(pre) if pre is not None else None
    _cabi.check(_cabi.lib().gb_nchw_to_cl(dout.data_ptr(), channels, C.byref(v), C.byref(pv) if pv is not None else None,
                                          1 if fp32 else 0, _stream()), "gb_nchw_to_cl")
    return dx


class MseConstFn(torch.autograd.Function):
    """mean((pred - target)^2) with the gradient produced in the same pass (adversarial_loss.py:60-62)."""

    @staticmethod
    def forward(ctx, pred, target):
        _require_cuda(pred, "prediction")
        pred = pred.contiguous().float()
        loss = torch.zeros((), dtype=torch.float32, device=pred.device)
        grad = torch.empty_like(pred) if ctx.needs_input_grad[0] else None
        _cabi.check(_cabi.lib().gb_mse_const(pred.data_ptr(), float(target), pred.numel(), loss.data_ptr(),
                                             grad.data_ptr() if grad is not None else None, _stream()), "gb_mse_const")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        return (grad * dloss if grad is not None else None), None


class L1Fn(torch.autograd.Function):
    """mean(|a - b|), gradient wrt a produced in the same pass (cyclegan_losses.py:64,73)."""

    @staticmethod
    def forward(ctx, a, b):
        _require_cuda(a, "L1 operand")
        a = a.contiguous().float()
        b = b.contiguous().float()
        loss = torch.zeros((), dtype=torch.float32, device=a.device)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        grad = torch.empty_like(a) if need else None
        _cabi.check(_cabi.lib().gb_l1(a.data_ptr(), b.data_ptr(), a.numel(), loss.data_ptr(),
                                      grad.data_ptr() if grad is not None else None, _stream()), "gb_l1")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        if grad is None:
            return None, None
        g = grad * dloss
        return (g if ctx.needs_input_grad[0] else None), (-g if ctx.needs_input_grad[1] else None)


class SsimFn(torch.autograd.Function):
    """SSIM distance of (x, y) mapped by (v + 1) / 2 (ganslate/nn/losses/cyclegan_losses.py:84-91 ->
    nn/losses/utils/ssim.py:64-99): one stencil kernel forward, one backward; gradient wrt x only (y is the real
    image in the reference's call)."""

    @staticmethod
    def forward(ctx, x, y, in_scale, in_shift, data_range):
        _require_cuda(x, "SSIM operand")
        if x.shape != y.shape or x.dim() not in (4, 5):
            raise ValueError("SSIM operands must be two NxCxHxW or NxCxDxHxW tensors of the same shape")
        xc, yc = x.contiguous().float(), y.detach().contiguous().float()
        H, W = xc.shape[-2:]
        planes = xc.numel() // (H * W)
        loss = torch.zeros((), dtype=torch.float32, device=x.device)
        _cabi.check(_cabi.lib().gb_ssim_fwd(xc.data_ptr(), yc.data_ptr(), planes, H, W, float(in_scale), float(in_shift),
                                            float(data_range), loss.data_ptr(), _stream()), "gb_ssim_fwd")
        ctx.save_for_backward(xc, yc)
        ctx.args = (planes, H, W, float(in_scale), float(in_shift), float(data_range))
        return loss

    @staticmethod
    def backward(ctx, dloss):
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None
        xc, yc = ctx.saved_tensors
        planes, H, W, sc, sh, dr = ctx.args
        dl = dloss.detach().contiguous().float().reshape(1)
        grad = torch.empty_like(xc)
        _cabi.check(_cabi.lib().gb_ssim_bwd(xc.data_ptr(), yc.data_ptr(), planes, H, W, sc, sh, dr, dl.data_ptr(),
                                            grad.data_ptr(), _stream()), "gb_ssim_bwd")
        return grad, None, None, None, None


class PatchNCEFn(torch.autograd.Function):
    """Per-row PatchNCE loss of (feat_q, feat_k) -- fused logits / mask / temperature / CE kernel
    (ganslate/nn/losses/cut_losses.py:14-43; feat_k is detached there, so only feat_q receives a gradient)."""

    @staticmethod
    def forward(ctx, feat_q, feat_k, batch, temperature):
        _require_cuda(feat_q, "PatchNCE features")
        q = feat_q.contiguous().float()
        k = feat_k.detach().contiguous().float()
        R, D = q.shape
        P = R // batch
        loss = torch.empty(R, dtype=torch.float32, device=q.device)
        probs = torch.empty((R, P + 1), dtype=torch.float32, device=q.device)
        _cabi.check(_cabi.lib().gb_patchnce_fwd(q.data_ptr(), k.data_ptr(), batch, P, D, float(temperature),
                                                loss.data_ptr(), probs.data_ptr(), _stream()), "gb_patchnce_fwd")
        ctx.save_for_backward(k, probs)
        ctx.cfg = (batch, P, D, float(temperature))
        return loss

    @staticmethod
    def backward(ctx, dloss):
        k, probs = ctx.saved_tensors
        batch, P, D, T = ctx.cfg
        dq = torch.empty((batch * P, D), dtype=torch.float32, device=k.device)
        dloss = dloss.contiguous().float()
        _cabi.check(_cabi.lib().gb_patchnce_bwd(k.data_ptr(), probs.data_ptr(), dloss.data_ptr(), batch, P, D, T,
                                                dq.data_ptr(), _stream()), "gb_patchnce_bwd")
        return dq, None, None, None


class PatchMlpFn(torch.autograd.Function):
    """CUT's FeaturePatchMLP for one feature (ganslate/nn/gans/unpaired/cut.py:262-276): gather the positions `ids`
    of feat (N, C, *spatial), Linear + ReLU + Linear, L2-normalise -- one fused launch forward (gb_patch_mlp_fwd), the
    row pass + two parameter-gradient launches backward (gb_patch_mlp_bwd).  fp32 FMA arithmetic, no library GEMM."""

    @staticmethod
    def forward(ctx, feat, ids, w1, b1, w2, b2):
        _require_cuda(feat, "feature map")
        f = feat.contiguous().float()
        N, Cc = f.shape[:2]
        F = f.numel() // (N * Cc)
        ids = ids.to(device=f.device, dtype=torch.int64).contiguous()
        P, nc = ids.numel(), w1.shape[0]
        assert w1.shape == (nc, Cc) and w2.shape == (nc, nc), (w1.shape, w2.shape, Cc)
        w1c, b1c, w2c, b2c = (t.detach().contiguous().float() for t in (w1, b1, w2, b2))
        R = N * P
        xg = torch.empty((R, Cc), dtype=torch.float32, device=f.device)
        h = torch.empty((R, nc), dtype=torch.float32, device=f.device)
        z = torch.empty((R, nc), dtype=torch.float32, device=f.device)
        y = torch.empty((R, nc), dtype=torch.float32, device=f.device)
        _call("patch_mlp", 0, "byte", "gb_patch_mlp_fwd", _cabi.lib().gb_patch_mlp_fwd, f.data_ptr(), ids.data_ptr(), N, Cc,
              F, P, w1c.data_ptr(), b1c.data_ptr(), w2c.data_ptr(), b2c.data_ptr(), nc, xg.data_ptr(), h.data_ptr(),
              z.data_ptr(), y.data_ptr(), _stream())
        ctx.save_for_backward(xg, h, z, ids, w1c, w2c)
        ctx.geom = (N, Cc, F, P, nc, tuple(feat.shape))
        return y

    @staticmethod
    def backward(ctx, dy):
        xg, h, z, ids, w1c, w2c = ctx.saved_tensors
        N, Cc, F, P, nc, fshape = ctx.geom
        dev = xg.device
        dy = dy.contiguous().float()
        R = N * P
        dz = torch.empty((R, nc), dtype=torch.float32, device=dev)
        dh = torch.empty((R, nc), dtype=torch.float32, device=dev)
        dfeat = torch.zeros(fshape, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        need_w = any(ctx.needs_input_grad[2:])
        dw1 = torch.empty_like(w1c) if need_w else None
        db1 = torch.empty(nc, dtype=torch.float32, device=dev) if need_w else None
        dw2 = torch.empty_like(w2c) if need_w else None
        db2 = torch.empty(nc, dtype=torch.float32, device=dev) if need_w else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        _call("patch_mlp", 0, "byte", "gb_patch_mlp_bwd", _cabi.lib().gb_patch_mlp_bwd, dy.data_ptr(), xg.data_ptr(),
              h.data_ptr(), z.data_ptr(), ids.data_ptr(), N, Cc, F, P, w1c.data_ptr(), w2c.data_ptr(), nc, dz.data_ptr(),
              dh.data_ptr(), ptr(dfeat), ptr(dw1), ptr(db1), ptr(dw2), ptr(db2), _stream())
        return dfeat, None, dw1, db1, dw2, db2
