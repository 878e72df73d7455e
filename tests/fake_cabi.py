"""TEST INFRASTRUCTURE ONLY -- a CPU restatement of the C ABI (include/ganslate_b200.h) at the pointer level, so that
the HOST logic above the ABI (ganslate_b200/ops.py, nn/layers.py, the network modules: tape construction, channel-slice
views, gradient routing, weight packing specs, reflection borders) can be exercised without a GPU.

It is NOT a fallback: nothing under ganslate_b200/ imports it, the product path rejects CPU tensors
(ops._require_cuda) and fails to import without the CUDA library.  tests/test_host_networks_cpu.py installs it with
monkeypatch for the duration of one test.  Each function follows the header's definition of the entry point (and,
where the header is terse, the kernel it documents: csrc/instnorm.cu, csrc/pack.cu, csrc/layout.cu, csrc/pad.cu),
written with plain torch CPU ops on tensors that alias the caller's memory.

What a test through this backend proves: the host side builds the right views / specs / tape for a network
(checked against the CPU oracle).  What it cannot prove: anything about the CUDA kernels themselves.
"""
import ctypes as C

import torch

from ganslate_b200 import _cabi
from ganslate_b200._cabi import ACT_LEAKY, ACT_NONE, ACT_PRELU, ACT_RELU, ACT_TANH

_DT = {2: torch.bfloat16, 4: torch.float32}


def _obj(a):
    """ctypes.byref(x) -> x; POINTER / raw struct pass through."""
    return a._obj if hasattr(a, "_obj") else a


def _flat(ptr, n, dtype):
    """1-D tensor of n elements aliasing the memory at address ptr."""
    if n <= 0:
        return torch.empty(0, dtype=dtype)
    esz = torch.empty(0, dtype=dtype).element_size()
    buf = (C.c_char * (n * esz)).from_address(int(ptr))
    return torch.frombuffer(buf, dtype=dtype, count=n)


def _strided(v, esz, border=False):
    """(N, D, H(+2p), W(+2p), C) tensor aliasing a gb_view; border=True includes the reflection border (negative
    coordinates relative to the interior origin)."""
    p = v.pad if border else 0
    H, W = v.H + 2 * p, v.W + 2 * p
    base = int(v.ptr) - p * (v.sy + v.sx) * esz
    if min(v.N, v.D, H, W, v.C) <= 0:
        return torch.empty((max(v.N, 0), max(v.D, 0), max(H, 0), max(W, 0), max(v.C, 0)), dtype=_DT[esz])
    n = (v.N - 1) * v.sn + (v.D - 1) * v.sz + (H - 1) * v.sy + (W - 1) * v.sx + v.C
    return _flat(base, n, _DT[esz]).as_strided((v.N, v.D, H, W, v.C), (v.sn, v.sz, v.sy, v.sx, 1))


def _act(v, act, slope):
    if act == ACT_RELU:
        return torch.clamp_min(v, 0.0)
    if act in (ACT_LEAKY, ACT_PRELU):
        return torch.where(v > 0, v, v * slope)
    if act == ACT_TANH:
        return torch.tanh(v)
    return v


def _mirror_lists(n, p):
    """For every interior index i of an axis of length n with reflection border p: the border coordinates that
    reflect onto it (PyTorch ReflectionPad: border -k mirrors k, border n-1+k mirrors n-1-k)."""
    out = [[] for _ in range(n)]
    for k in range(1, p + 1):
        out[k].append(-k)
        out[n - 1 - k].append(n - 1 + k)
    return out


def _fold_border(t, v, p):
    """t: (N, D, H+2p, W+2p, C) gradient on the padded domain -> (N, D, H, W, C): interior + mirrored border."""
    if p == 0:
        return t.clone()
    H, W = v.H, v.W
    rows = torch.zeros((t.shape[0], t.shape[1], H, W + 2 * p, t.shape[4]), dtype=t.dtype)
    rows += t[:, :, p:p + H]
    for i, lst in enumerate(_mirror_lists(H, p)):
        for b in lst:
            rows[:, :, i] += t[:, :, b + p]
    out = rows[:, :, :, p:p + W].clone()
    for i, lst in enumerate(_mirror_lists(W, p)):
        for b in lst:
            out[:, :, :, i] += rows[:, :, :, b + p]
    return out


def _write_reflect(dst_b, val, v):
    """dst_b: bordered alias (N, D, H+2p, W+2p, C); val: interior values -> interior + reflection border."""
    p = v.pad
    if p == 0:
        dst_b.copy_(val)
        return
    H, W = v.H, v.W
    full = torch.nn.functional.pad(val.permute(0, 1, 4, 2, 3).reshape(-1, val.shape[4], H, W).float(), (p, p, p, p),
                                   mode="reflect")
    full = full.reshape(val.shape[0], val.shape[1], val.shape[4], H + 2 * p, W + 2 * p).permute(0, 1, 3, 4, 2)
    dst_b.copy_(full.to(dst_b.dtype))


class FakeLib:
    """Drop-in for ctypes.CDLL(libganslate_b200.so) on CPU memory (see module docstring)."""

    def __init__(self):
        self.knobs = [0] * 32
        self.launches = 0
        self.calls = {}

    def _count(self, name):
        self.launches += 1
        self.calls[name] = self.calls.get(name, 0) + 1

    # ------------------------------------------------------------------ plumbing
    def gb_version(self):
        return 100

    def gb_last_error(self):
        return b"fake backend"

    def gb_launch_count(self):
        return self.launches

    def gb_debug_knob(self, k, v):
        old, self.knobs[k] = self.knobs[k], v
        return old

    def gb_tma_window_supported(self):
        return 1

    # ------------------------------------------------------------------ weights
    def _pack_one(self, p):
        nsrc = 0
        tids = [p.tap_id[i] for i in range(_cabi.GB_MAX_TAPS)]
        for cls in range(p.nclass):
            kpad, ntaps, tb = p.kpad[cls], p.ntaps[cls], p.tap_begin[cls]
            dst = _flat(int(p.dst) + 2 * p.w_offset[cls], p.rows_pad * kpad, torch.bfloat16).view(p.rows_pad, kpad)
            dst.zero_()
            tmax = max([t for t in tids[tb:tb + ntaps] if t >= 0], default=-1)
            if tmax < 0:
                continue
            nsrc = (p.rows - 1) * p.sn + (p.chans - 1) * p.sc + tmax * p.st + 1
            src = _flat(p.src, nsrc, torch.float32)
            n = torch.arange(p.rows).view(-1, 1)
            c = torch.arange(p.chans).view(1, -1)
            for tl in range(ntaps):
                t = tids[tb + tl]
                if t < 0 or tl * p.chans_pad >= kpad:
                    continue
                vals = src[n * p.sn + c * p.sc + t * p.st]
                dst[:p.rows, tl * p.chans_pad:tl * p.chans_pad + p.chans] = vals.to(torch.bfloat16)

    def gb_pack_weights(self, p, stream):
        self._count("gb_pack_weights")
        self._pack_one(_obj(p))
        return 0

    def gb_pack_weights_multi(self, table, count, max_elems, stream):
        self._count("gb_pack_weights_multi")
        arr = (_cabi.PackParams * count).from_address(int(table))
        for i in range(count):
            self._pack_one(arr[i])
        return 0

    def _unpack(self, dw, dst, dsr, dsc, dst_t, rows, chans, chans_pad, ntaps, kpad, accumulate=0):
        w = _flat(dw, rows * kpad, torch.float32).view(rows, kpad)
        r = torch.arange(rows).view(-1, 1, 1)
        c = torch.arange(chans).view(1, -1, 1)
        t = torch.arange(ntaps).view(1, 1, -1)
        idx = r * dsr + c * dsc + t * dst_t
        lo, hi = int(idx.min()), int(idx.max())
        out = _flat(int(dst) + 4 * lo, hi - lo + 1, torch.float32)
        vals = w[:, :ntaps * chans_pad].reshape(rows, ntaps, chans_pad)[:, :, :chans].permute(0, 2, 1)
        flat_idx = (idx - lo).reshape(-1)
        if accumulate:
            out[flat_idx] += vals.reshape(-1)
        else:
            out[flat_idx] = vals.reshape(-1)

    def gb_unpack_wgrad(self, dw, dst, dsr, dsc, dst_t, rows, chans, chans_pad, ntaps, kpad, stream):
        self._count("gb_unpack_wgrad")
        self._unpack(dw, dst, dsr, dsc, dst_t, rows, chans, chans_pad, ntaps, kpad)
        return 0

    def gb_unpack_wgrad_multi(self, b, stream):
        self._count("gb_unpack_wgrad_multi")
        b = _obj(b)
        for i in range(b.count):
            it = b.item[i]
            self._unpack(it.dw, it.dst, it.dsr, it.dsc, it.dst_t, it.rows, it.chans, it.chans_pad, it.ntaps, it.kpad,
                         it.accumulate)
        return 0

    def gb_colsum(self, v, out, stream):
        self._count("gb_colsum")
        v = _obj(v)
        t = _strided(v, 2).float()
        _flat(out, v.C, torch.float32).copy_(t.sum(dim=(0, 1, 2, 3)))
        return 0

    # ------------------------------------------------------------------ convolutions
    @staticmethod
    def _gather(x, ext, q_ext, mul, tap, c_valid):
        """x: (N, D, H, W, C) float alias (possibly with overlapping strides) -> (N, qz, qy, qx, C) gathered at
        q * mul + tap, zero outside [0, ext) and for channels >= c_valid."""
        idx, ok = [], []
        for d in range(3):
            i = torch.arange(q_ext[d]) * mul[d] + tap[d]
            ok.append((i >= 0) & (i < ext[d]))
            idx.append(i.clamp(0, max(ext[d] - 1, 0)))
        g = x[:, idx[0]][:, :, idx[1]][:, :, :, idx[2]].float()
        m = (ok[0].view(-1, 1, 1) & ok[1].view(1, -1, 1) & ok[2].view(1, 1, -1)).view(1, *q_ext, 1)
        g = g * m
        if c_valid:
            g[..., c_valid:] = 0
        return g

    def gb_conv_data(self, p, stream):
        self._count("gb_conv_data")
        p = _obj(p)
        assert p.inp.pad == 0, "gb_conv_data reads a plain view"
        vin, vout = p.inp, p.out
        x = _strided(vin, 2)
        if p.in_c_valid:  # window view: do not touch memory behind the last real channel of the last pixel
            x = _strided_window(vin, p.in_c_valid)
        out = _strided(vout, 4 if p.out_fp32 else 2)
        ext_in, ext_out = (vin.D, vin.H, vin.W), (vout.D, vout.H, vout.W)
        bias = _flat(p.bias, p.ncols, torch.float32) if p.bias else None
        stats = _flat(p.stats, vout.N * vout.C * 2, torch.float32).view(vout.N, vout.C, 2) if p.stats else None
        in_mul, out_mul = list(p.in_mul), list(p.out_mul)
        for ci in range(p.nclass):
            cc = p.cls[ci]
            off = list(cc.off)
            q_ext = [max(0, -(-(ext_out[d] - off[d]) // out_mul[d])) for d in range(3)]
            if min(q_ext) == 0:
                continue
            w = _flat(int(p.wpacked) + 2 * cc.w_offset, p.npad * cc.kpad, torch.bfloat16).view(p.npad, cc.kpad).float()
            acc = torch.zeros((vin.N, *q_ext, p.npad))
            for tl in range(cc.ntaps):
                tap = [int(p.taps[cc.tap_begin + tl][d]) for d in range(3)]
                g = self._gather(x, ext_in, q_ext, in_mul, tap, p.in_c_valid)
                acc += g @ w[:, tl * vin.C:(tl + 1) * vin.C].T
            ncol = min(vout.C, p.npad)
            val = acc[..., :ncol]
            if bias is not None:
                val[..., :p.ncols] += bias
            val = _act(val, p.act, p.act_slope)
            sl = tuple(slice(off[d], off[d] + q_ext[d] * out_mul[d], out_mul[d]) for d in range(3))
            tgt = out[:, sl[0], sl[1], sl[2], :ncol]
            if p.out_fp32:
                tgt.copy_(tgt + val if p.accumulate else val)
            else:
                vb = val.to(torch.bfloat16)
                tgt.copy_(vb)
                if stats is not None:
                    f = vb.float()[..., :p.ncols]
                    stats[:, :p.ncols, 0] += f.sum(dim=(1, 2, 3))
                    stats[:, :p.ncols, 1] += (f * f).sum(dim=(1, 2, 3))
        self.knobs[15] = 2
        return 0

    def gb_conv_wgrad(self, p, stream):
        self._count("gb_conv_wgrad")
        p = _obj(p)
        vp, vg = p.plain, p.gathered
        plain = _strided(vp, 2).float()
        gath = _strided_window(vg, p.gathered_c_valid) if p.gathered_c_valid else _strided(vg, 2)
        rows_alloc = (p.rows + 127) // 128 * 128
        dw = _flat(p.dw, rows_alloc * p.kpad, torch.float32).view(rows_alloc, p.kpad)
        q_ext, mul = (vp.D, vp.H, vp.W), list(p.mul)
        r = min(p.rows, vp.C)
        P2 = plain[..., :r].reshape(-1, r)
        for t in range(p.ntaps):
            if t * vg.C >= p.kpad:
                break
            tap = [int(p.taps[t][d]) for d in range(3)]
            g = self._gather(gath, (vg.D, vg.H, vg.W), q_ext, mul, tap, p.gathered_c_valid).reshape(-1, vg.C)
            ncol = min(vg.C, p.kpad - t * vg.C)
            dw[:r, t * vg.C:t * vg.C + ncol] += (P2.T @ g)[:, :ncol]
        self.knobs[14] = 1
        return 0

    # ------------------------------------------------------------------ InstanceNorm family
    def gb_in_stats(self, v, stats, stream):
        self._count("gb_in_stats")
        v = _obj(v)
        x = _strided(v, 2).float()
        st = _flat(stats, v.N * v.C * 2, torch.float32).view(v.N, v.C, 2)
        st[:, :, 0] += x.sum(dim=(1, 2, 3))
        st[:, :, 1] += (x * x).sum(dim=(1, 2, 3))
        return 0

    @staticmethod
    def _moments(stats_ptr, v, eps):
        P = v.D * v.H * v.W
        st = _flat(stats_ptr, v.N * v.C * 2, torch.float32).view(v.N, v.C, 2)
        mean = st[:, :, 0] / P
        var = torch.clamp_min(st[:, :, 1] / P - mean * mean, 0.0)
        rstd = torch.rsqrt(var + eps)
        return mean.view(v.N, 1, 1, 1, v.C), rstd.view(v.N, 1, 1, 1, v.C)

    def gb_in_fwd(self, p, stream):
        self._count("gb_in_fwd")
        p = _obj(p)
        v = p.x
        x = _strided(v, 2).float()
        if p.stats:
            mean, rstd = self._moments(p.stats, v, p.eps)
            val = (x - mean) * rstd
        else:
            val = x
        slope = _flat(p.prelu, v.C, torch.float32).view(1, 1, 1, 1, -1) if p.act == ACT_PRELU else p.act_slope
        res = _strided(p.res, 2).float() if p.res.ptr else None
        oscale = 1.0 if p.out_scale == 0.0 else p.out_scale
        if res is not None and p.res_before_act:
            val = val + res
        val = _act(val, p.act, slope) * oscale
        if res is not None and not p.res_before_act:
            val = val + res
        _write_reflect(_strided(p.y, 2, border=True), val.to(torch.bfloat16), p.y)
        return 0

    def gb_in_bwd(self, p, stream):
        self._count("gb_in_bwd")
        p = _obj(p)
        v = p.x
        norm = bool(p.stats)
        # incoming gradient: dy_a (plain) + fold(dy_b on the padded domain)
        g = torch.zeros((v.N, v.D, v.H, v.W, v.C))
        if p.dy_b.ptr:
            g = g + _fold_border(_strided(p.dy_b, 4, border=True), p.dy_b, p.dy_b.pad)
        if p.dy_a.ptr:
            g = g + _strided(p.dy_a, 4)
        rba = bool(p.res_before_act) and bool(p.res.ptr)
        dy_sum = _strided(p.dy_sum, 4) if p.dy_sum.ptr else None
        if dy_sum is not None and not rba:  # residual added after the activation: unmasked total
            dy_sum.copy_(dy_sum + g if p.dy_sum_acc else g)
        oscale = 1.0 if p.out_scale == 0.0 else p.out_scale
        g = g * oscale
        slope = _flat(p.prelu, v.C, torch.float32).view(1, 1, 1, 1, -1) if p.act == ACT_PRELU else p.act_slope
        x = _strided(v, 2).float()
        res = _strided(p.res, 2).float() if rba else None
        xh = None
        dprelu = None
        if norm:
            mean, rstd = self._moments(p.stats, v, p.eps)
            xh = (x - mean) * rstd
            pre = xh + res if rba else xh
        elif p.act != ACT_NONE:
            if p.y.ptr:
                pre = _strided(p.y, 2).float()      # derivative from the forward output
            else:
                pre = x + res if rba else x
        if p.act in (ACT_RELU, ACT_LEAKY, ACT_PRELU) and (norm or p.act != ACT_NONE):
            neg = ~(pre > 0)
            if p.act == ACT_PRELU and p.dprelu:
                dprelu = (g * pre * neg).sum(dim=(0, 1, 2, 3))
            g = torch.where(neg, g * (0.0 if p.act == ACT_RELU else slope), g)
        elif p.act == ACT_TANH:
            th = pre if (p.y.ptr and not norm) else torch.tanh(pre)
            g = g * (1 - th * th)
        if dy_sum is not None and rba:  # residual added before the activation: masked gradient
            dy_sum.copy_(dy_sum + g if p.dy_sum_acc else g)
        if norm:
            P = v.D * v.H * v.W
            m1 = g.sum(dim=(1, 2, 3), keepdim=True) / P
            m2 = (g * xh).sum(dim=(1, 2, 3), keepdim=True) / P
            d = rstd * (g - m1 - xh * m2)
        else:
            d = g
        if dprelu is not None:
            _flat(p.dprelu, v.C, torch.float32).add_(dprelu)
        if p.dx_fp32_acc:
            dx = _strided(p.dx, 4)
            dx.copy_(dx + d)
        else:
            _strided(p.dx, 2).copy_(d.to(torch.bfloat16))
            if p.dbias:
                _flat(p.dbias, v.C, torch.float32).add_(d.sum(dim=(0, 1, 2, 3)))
        return 0

    # ------------------------------------------------------------------ losses
    def gb_mse_const(self, pred, target, n, loss, grad, stream):
        self._count("gb_mse_const")
        d = _flat(pred, n, torch.float32) - target
        _flat(loss, 1, torch.float32).add_((d * d).sum() / n)
        if grad:
            _flat(grad, n, torch.float32).copy_(2.0 * d / n)
        return 0

    def gb_l1(self, a, b, n, loss, grad, stream):
        self._count("gb_l1")
        d = _flat(a, n, torch.float32) - _flat(b, n, torch.float32)
        _flat(loss, 1, torch.float32).add_(d.abs().sum() / n)
        if grad:
            _flat(grad, n, torch.float32).copy_(torch.sign(d) / n)
        return 0

    def _nce_logits(self, q, k, B, P, D, T):
        qq, kk = _flat(q, B * P * D, torch.float32).view(B, P, D), _flat(k, B * P * D, torch.float32).view(B, P, D)
        l_pos = (qq * kk).sum(-1, keepdim=True)                       # own key
        l_neg = torch.bmm(qq, kk.transpose(1, 2))                      # keys of the same image
        l_neg = l_neg.masked_fill(torch.eye(P, dtype=torch.bool)[None], -10.0)
        return torch.cat([l_pos, l_neg], dim=2).view(B * P, P + 1) / T, kk

    def gb_patchnce_fwd(self, q, k, B, P, D, T, loss, probs, stream):
        self._count("gb_patchnce_fwd")
        lg, _ = self._nce_logits(q, k, B, P, D, T)
        lse = torch.logsumexp(lg, dim=1)
        _flat(loss, B * P, torch.float32).copy_(lse - lg[:, 0])
        if probs:
            _flat(probs, B * P * (P + 1), torch.float32).view(B * P, P + 1).copy_(torch.exp(lg - lse[:, None]))
        return 0

    def gb_patchnce_bwd(self, k, probs, dloss, B, P, D, T, dq, stream):
        self._count("gb_patchnce_bwd")
        kk = _flat(k, B * P * D, torch.float32).view(B, P, D)
        c = _flat(probs, B * P * (P + 1), torch.float32).view(B, P, P + 1).clone()
        c[:, :, 0] -= 1.0
        idx = torch.arange(P)
        c[:, idx, 1 + idx] = 0.0                                       # the masked diagonal is a constant
        c = c * (_flat(dloss, B * P, torch.float32).view(B, P, 1) / T)
        out = c[:, :, :1] * kk + torch.bmm(c[:, :, 1:], kk)
        _flat(dq, B * P * D, torch.float32).view(B, P, D).copy_(out)
        return 0

    def gb_patch_mlp_fwd(self, feat, ids, N, C, F, P, W1, b1, W2, b2, nc, xg, h, z, y, stream):
        self._count("gb_patch_mlp_fwd")
        f = _flat(feat, N * C * F, torch.float32).view(N, C, F)
        idx = _flat(ids, P, torch.int64)
        x = f[:, :, idx].permute(0, 2, 1).reshape(N * P, C)
        w1, w2 = _flat(W1, nc * C, torch.float32).view(nc, C), _flat(W2, nc * nc, torch.float32).view(nc, nc)
        hh = torch.relu(x @ w1.t() + _flat(b1, nc, torch.float32))
        zz = hh @ w2.t() + _flat(b2, nc, torch.float32)
        yy = zz / (zz.pow(2).sum(1, keepdim=True).sqrt() + 1e-7)
        for dst, src, n in ((xg, x, N * P * C), (h, hh, N * P * nc), (z, zz, N * P * nc), (y, yy, N * P * nc)):
            _flat(dst, n, torch.float32).copy_(src.reshape(-1))
        return 0

    def gb_patch_mlp_bwd(self, dy, xg, h, z, ids, N, C, F, P, W1, W2, nc, dz, dh, dfeat, dW1, db1, dW2, db2, stream):
        self._count("gb_patch_mlp_bwd")
        R = N * P
        g = _flat(dy, R * nc, torch.float32).view(R, nc)
        x, hh, zz = (_flat(xg, R * C, torch.float32).view(R, C), _flat(h, R * nc, torch.float32).view(R, nc),
                     _flat(z, R * nc, torch.float32).view(R, nc))
        w1, w2 = _flat(W1, nc * C, torch.float32).view(nc, C), _flat(W2, nc * nc, torch.float32).view(nc, nc)
        nrm = zz.pow(2).sum(1, keepdim=True).sqrt()
        inv = 1.0 / (nrm + 1e-7)
        gz = g * inv - zz * ((g * zz).sum(1, keepdim=True) * inv * inv / nrm)
        gh = (gz @ w2) * (hh > 0)
        _flat(dz, R * nc, torch.float32).copy_(gz.reshape(-1))
        _flat(dh, R * nc, torch.float32).copy_(gh.reshape(-1))
        if dfeat:
            idx = _flat(ids, P, torch.int64)
            d = _flat(dfeat, N * C * F, torch.float32).view(N, C, F)
            d[:, :, idx] = (gh @ w1).view(N, P, C).permute(0, 2, 1)
        if dW2:
            _flat(dW2, nc * nc, torch.float32).copy_((gz.t() @ hh).reshape(-1))
            _flat(db2, nc, torch.float32).copy_(gz.sum(0))
        if dW1:
            _flat(dW1, nc * C, torch.float32).copy_((gh.t() @ x).reshape(-1))
            _flat(db1, nc, torch.float32).copy_(gh.sum(0))
        return 0

    # ------------------------------------------------------------------ layout / padding
    def gb_nchw_to_cl(self, src, Cc, dst, pre, dst_fp32, stream):
        self._count("gb_nchw_to_cl")
        v = _obj(dst)
        P = v.D * v.H * v.W
        s = _flat(src, v.N * Cc * P, torch.float32).view(v.N, Cc, v.D, v.H, v.W).permute(0, 2, 3, 4, 1)
        val = torch.zeros((v.N, v.D, v.H, v.W, v.C))
        val[..., :Cc] = s
        if pre:
            pv = _obj(pre)
            if pv.ptr:
                th = torch.tanh(_strided(pv, 2).float())
                val = val * (1 - th * th)
        tgt = _strided(v, 4 if dst_fp32 else 2, border=True)
        _write_reflect(tgt, val if dst_fp32 else val.to(torch.bfloat16), v)
        return 0

    def gb_cl_to_nchw(self, src, dst, Cc, fold, act, src_fp32, stream):
        self._count("gb_cl_to_nchw")
        v = _obj(src)
        esz = 4 if src_fp32 else 2
        if fold and v.pad > 0:
            t = _fold_border(_strided(v, esz, border=True).float(), v, v.pad)
        else:
            t = _strided(v, esz).float()
        t = t[..., :Cc]
        if act == ACT_TANH:
            t = torch.tanh(t)
        P = v.D * v.H * v.W
        _flat(dst, v.N * Cc * P, torch.float32).view(v.N, Cc, v.D, v.H, v.W).copy_(t.permute(0, 4, 1, 2, 3))
        return 0

    def gb_replicate_pad_fwd(self, src, dst, pz, py, px, stream):
        self._count("gb_replicate_pad_fwd")
        s, d = _obj(src), _obj(dst)
        x = _strided(s, 2).float().permute(0, 4, 1, 2, 3)
        y = torch.nn.functional.pad(x, (px, px, py, py, pz, pz), mode="replicate")
        _strided(d, 2).copy_(y.permute(0, 2, 3, 4, 1).to(torch.bfloat16))
        return 0

    def gb_replicate_pad_bwd(self, ddst, dsrc, pz, py, px, stream):
        self._count("gb_replicate_pad_bwd")
        g, d = _obj(ddst), _obj(dsrc)
        with torch.enable_grad():  # called from inside an autograd.Function's backward
            x = torch.zeros((d.N, d.C, d.D, d.H, d.W), requires_grad=True)
            y = torch.nn.functional.pad(x, (px, px, py, py, pz, pz), mode="replicate")
            (gx,) = torch.autograd.grad(y, x, _strided(g, 4).permute(0, 4, 1, 2, 3).contiguous())
        tgt = _strided(d, 4)
        tgt.copy_(tgt + gx.permute(0, 2, 3, 4, 1))
        return 0


def _strided_window(v, c_valid):
    """Pixel-window view (C = 64 spans 8 pixels of an 8-channel tensor, sx = 8): the last window of the allocation
    would alias memory behind the buffer, so materialise a copy that only reads the c_valid leading channels."""
    n = (v.N - 1) * v.sn + (v.D - 1) * v.sz + (v.H - 1) * v.sy + (v.W - 1) * v.sx + c_valid
    flat = _flat(v.ptr, n, torch.bfloat16)
    out = torch.zeros((v.N, v.D, v.H, v.W, v.C), dtype=torch.bfloat16)
    out[..., :c_valid] = flat.as_strided((v.N, v.D, v.H, v.W, c_valid), (v.sn, v.sz, v.sy, v.sx, 1))
    return out


def install(monkeypatch):
    """Route ganslate_b200's host code to the fake backend for one test (pytest monkeypatch: undone afterwards)."""
    from ganslate_b200 import ops
    lib = FakeLib()
    monkeypatch.setattr(_cabi, "_lib", lib)
    monkeypatch.setattr(_cabi, "lib", lambda: lib)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "_require_cuda", lambda t, what: None)
    return lib
