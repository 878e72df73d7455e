"""Teacher-forced layer-by-layer parity of a whole network against the bf16-point CPU oracle.

Why: two bf16 pipelines that differ only in fp32 summation order round a few elements of every layer differently, and
every later rounding turns a difference d << ulp into sqrt(ulp * d): after ~5 layers the two are as far apart as either
is from fp32 (measured: tests/parity_matrix.py, BASELINE.md "Tolerances").  End-to-end agreement of DEEP networks is
therefore bounded by the precision choice, not by the kernels.  What CAN be asserted tightly is that every fused launch
of the real network, at the real shape, on the real data, computes what the reference computes FROM THE SAME INPUTS:

  forward   after each fused step (convolution [+bias, activation] / InstanceNorm + activation + residual + border) the
            B200 buffer is compared with the oracle's tensor at the same point and then OVERWRITTEN with it
            (statistics of a forced raw convolution output are recomputed from the forced values);
  backward  before each step's backward runs, the gradient that has arrived at its output is compared with the
            oracle's gradient at that tensor and overwritten with it;
  so every kernel launch sees oracle inputs, and every per-parameter gradient is a function of oracle tensors only.

One effect survives forcing: the derivative of ReLU / LeakyReLU is taken at IN(x), and an element whose normalised
value is within summation-order noise of ZERO can get the other branch -- an isolated element of a layer gradient that
is off by the whole gradient value (seen: 1 - 20 elements of 4 M).  The per-tensor statistics therefore are relative
L2, max-relative error AND the fraction of elements further than 1e-2 x max|ref| from the reference.

The oracle is oracle/torch_oracle.py::forward_bf16_points (TRACE_LIVE) -- test infrastructure.  Works for the
sequence networks (Resnet2D, PatchGAN2D/3D): those are what BASELINE configs 1-3 run."""
import torch
import torch.nn.functional as F

from ganslate_b200.nn import layers
from oracle import torch_oracle as O


def _interior(buf, fp32=False):
    """(N, C, [D,] H, W) fp32 view of a Buf's interior (logical channels)."""
    t, p = buf.st.t, buf.st.pad
    x = t[:, :, p:t.shape[2] - p, p:t.shape[3] - p, buf.c0:buf.c0 + buf.channels]
    return x


def _to_ref_layout(x5, is_3d):
    x = x5.permute(0, 4, 1, 2, 3)
    return x if is_3d else x[:, :, 0]


def _from_ref_layout(t, is_3d):
    return (t if is_3d else t.unsqueeze(2)).permute(0, 2, 3, 4, 1)


OUTLIER = 1e-2   # an element is an outlier when |ours - ref| > OUTLIER * max|ref|


def _errs(a, b):
    """(relative L2, max-relative error, fraction of outlier elements)."""
    a, b = a.float().cpu(), b.float().cpu()
    d = (a - b).abs()
    scale = b.abs().max() + 1e-30
    return ((a - b).norm() / (b.norm() + 1e-30)).item(), (d.max() / scale).item(), (d > OUTLIER * scale).float().mean().item()


def _fold(g, p):
    """Fold the reflection border of a (N, D, H+2p, W+2p, C) gradient onto its interior (transpose of ReflectionPad)."""
    if p == 0:
        return g
    g = g.clone()
    H, W = g.shape[2] - 2 * p, g.shape[3] - 2 * p
    for i in range(p):
        g[:, :, 2 * p - i] += g[:, :, i]                      # top rows mirror about row p
        g[:, :, H + p - 2 - (p - 1 - i)] += g[:, :, H + p + (p - 1 - i)]
    g = g[:, :, p:H + p]
    for i in range(p):
        g[:, :, :, 2 * p - i] += g[:, :, :, i]
        g[:, :, :, W + p - 2 - (p - 1 - i)] += g[:, :, :, W + p + (p - 1 - i)]
    return g[:, :, :, p:W + p]


class Forcer:
    """Context manager that patches layers.step_conv / step_norm_act for ONE network evaluation."""

    def __init__(self, trace, force=True, fp32=False):
        self.trace = trace          # [(kind, live oracle tensor)] from O.TRACE with TRACE_LIVE
        self.force = force
        self.fp32 = fp32            # fp32 validation mode: fp32 buffers, no bf16 rounding of the reference gradients
        self.i = 0
        self.fwd, self.bwd = [], []  # (index, kind, shape, rel_l2, max_rel, outlier fraction)

    def __enter__(self):
        self._sc, self._sn = layers.step_conv, layers.step_norm_act
        layers.step_conv = lambda tape, b, m, *a, **k: self._after(tape, self._sc(tape, b, m, *a, **k))
        layers.step_norm_act = lambda tape, *a, **k: self._after(tape, self._sn(tape, *a, **k))
        return self

    def __exit__(self, *exc):
        layers.step_conv, layers.step_norm_act = self._sc, self._sn

    def _after(self, tape, out):
        idx = self.i
        self.i += 1
        kind, ref = self.trace[idx]
        # (a network's last convolution without activation is "act" for the oracle's trace and a raw buffer here:
        #  what matters below is how THIS path stores the gradient of the buffer, bf16 d_raw or fp32)
        kind = "raw" if out.raw else "act"
        mine = _to_ref_layout(_interior(out), out.is_3d)
        assert tuple(mine.shape) == tuple(ref.shape), (idx, mine.shape, ref.shape)
        self.fwd.append((idx, kind, tuple(ref.shape)) + _errs(mine, ref.detach()))
        dev = out.st.t.device
        if self.force:
            r = ref.detach()
            p = out.st.pad
            if p:
                r = F.pad(r, (p, p, p, p), mode="reflect")
            r5 = _from_ref_layout(r, out.is_3d).to(dev)
            out.st.t[..., out.c0:out.c0 + out.channels] = r5.to(out.st.t.dtype)
            if out.stats is not None:
                v = ref.detach().double()
                dims = tuple(range(2, v.dim()))
                st = torch.stack([v.sum(dims), (v * v).sum(dims)], dim=-1).float()   # (N, C, 2)
                out.stats.zero_()
                out.stats[:, :out.channels] = st.to(dev)
        if tape is not None:
            inner = tape.steps[-1]

            def bwd():
                g = out.st.grad
                if g is not None and ref.grad is not None:
                    rg = ref.grad.detach()
                    if kind == "raw" and not self.fp32:
                        rg = rg.to(torch.bfloat16).float()       # the oracle rounds d_raw when it propagates it
                    p = out.st.pad
                    mine = _to_ref_layout(_fold(g.float(), p)[..., out.c0:out.c0 + out.channels], out.is_3d)
                    self.bwd.append((idx, kind, tuple(rg.shape)) + _errs(mine, rg))
                    if self.force:
                        g.zero_()
                        r5 = _from_ref_layout(rg, out.is_3d).to(dev).to(g.dtype)
                        g[:, :, p:g.shape[2] - p, p:g.shape[3] - p, out.c0:out.c0 + out.channels] = r5
                inner()

            tape.steps[-1] = bwd
        return out


def forced_network_parity(ours, ref, x, dy=None, force=True, fp32=False):
    """Evaluate `ours` (cuda / fake backend) and `ref` (oracle module, same weights) on x with teacher forcing.
    fp32: the fp32 validation mode (ops.FP32_MODE must be on) against the oracle walked WITHOUT rounding.
    Returns dict(fwd=[...], bwd=[...], out=(rel_l2, max_rel), dx=..., params={name: (rel_l2, max_rel, |ref|max)})."""
    O.TRACE, O.TRACE_LIVE = [], True
    O.ROUND_BF16 = not fp32
    try:
        xr = x.clone().requires_grad_(True)
        yr = O.forward_bf16_points(ref, xr)
        trace = O.TRACE
    finally:
        O.TRACE, O.TRACE_LIVE = None, False
    if dy is None:
        g = torch.Generator().manual_seed(1234)
        dy = torch.randn(yr.shape, generator=g)
    ref.zero_grad()
    try:
        yr.backward(dy)
    finally:
        O.ROUND_BF16 = True
    dev = next(ours.parameters()).device
    xo = x.clone().to(dev).requires_grad_(True)
    with Forcer(trace, force, fp32) as f:
        yo = ours(xo)
        assert f.i == len(trace), (f.i, len(trace))
        ours.zero_grad()
        yo.backward(dy.to(dev))
    rep = dict(fwd=f.fwd, bwd=f.bwd, out=_errs(yo.detach(), yr.detach()), dx=_errs(xo.grad, xr.grad), params={})
    po = dict(ours.named_parameters())
    for k, p in ref.named_parameters():
        if p.grad is not None:
            rep["params"][k] = _errs(po[k].grad, p.grad)[:2] + (p.grad.abs().max().item(),)
    return rep


def summarize(rep):
    w = {k: v for k, v in rep["params"].items() if k.endswith("weight")}
    return dict(fwd_max_rel=max(v[4] for v in rep["fwd"]), fwd_rel_l2=max(v[3] for v in rep["fwd"]),
                bwd_max_rel=max(v[4] for v in rep["bwd"]), bwd_rel_l2=max(v[3] for v in rep["bwd"]),
                fwd_outlier_frac=max(v[5] for v in rep["fwd"]), bwd_outlier_frac=max(v[5] for v in rep["bwd"]),
                out=rep["out"], dx=rep["dx"], wgrad_max_rel=max(v[1] for v in w.values()),
                wgrad_rel_l2=max(v[0] for v in w.values()))
