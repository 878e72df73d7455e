"""Times single fused InstanceNorm-backward launches (gb_in_bwd) on the layer shapes of a CycleGAN step under the
kernel variants selectable with bring-up knobs, with CUDA events on the launch stream, cold (L2 flushed before every
launch) and hot (the tensors were just touched: the in-step situation, where the gradient was written by the
preceding data-gradient launch).  Not a bench value: the A/B tool for choosing the default variant (copy the table
into profiles/ by hand).  Every variant's dx / bias gradient / residual gradient is checked against variant 0.

    python tools/in_microbench.py [batch]

variants: knob 24 = 1 / 2 on-chip cluster kernel (instnorm_v3.cu: every byte read once; 2 = half-size stash, two CTAs
per SM; maps of more than 8192 pixels are declined and fall through to the knob-22 choice); knob 22 = 0 first generation (instnorm_fast.cu, 4 channels per thread), 1 / 2 second generation
(instnorm_v2.cu, 8 channels per thread, 4 / 2 pixels in flight); knob 6 = 1 two launches instead of one launch around
a grid barrier.  GB/s = algorithmic bytes (fp32 gradient read once + bf16 x read once + bf16 dx written, + 8 B per
element read-modify-write of the residual gradient) / time; the kernels read the gradient and x twice, so the
traffic the memory system sees is higher.
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ganslate_b200 import _cabi, ops
from ganslate_b200._cabi import ACT_LEAKY, ACT_NONE, ACT_RELU

dev = "cuda"
VARIANTS = [("gen1 fused", {22: 0, 6: 2}), ("gen1 2-launch", {22: 0, 6: 1}), ("lean fused", {22: 4, 6: 2}),
            ("lean 2-launch", {22: 4, 6: 1}), ("gen2 U4 fused", {22: 1, 6: 2}),
            ("gen2 U4 2-launch", {22: 1, 6: 1}), ("gen2 U2 fused", {22: 2, 6: 2}), ("on-chip cluster", {24: 1}),
            ("on-chip half stash", {24: 2})]


def make(name, N, Cc, H, W, act, gpad, res):
    x = (torch.randn(N, 1, H, W, Cc, device=dev) * 1.3 + 0.4).to(torch.bfloat16)
    dy = torch.randn(N, 1, H + 2 * gpad, W + 2 * gpad, Cc, device=dev)
    xf = x.float()
    stats = torch.stack([xf.sum(dim=(1, 2, 3)), (xf * xf).sum(dim=(1, 2, 3))], dim=-1).contiguous()
    E = N * H * W * Cc
    return dict(name=name, x=x, dy=dy, stats=stats, act=act, slope=0.2 if act == ACT_LEAKY else 0.0, gpad=gpad, res=res,
                bytes=E * (8 + (8 if res else 0)), shape=(N, Cc, H, W))


def launch(L, out):
    p = _cabi.InBwdParams()
    p.x, p.dy_b, p.dx = ops.make_view(L["x"]), ops.make_view(L["dy"], L["gpad"]), ops.make_view(out["dx"])
    if L["res"]:
        p.dy_sum, p.dy_sum_acc = ops.make_view(out["dsum"]), 1
    p.stats, p.bstats, p.dbias = L["stats"].data_ptr(), out["bstats"].data_ptr(), out["dbias"].data_ptr()
    p.eps, p.act, p.act_slope = 1e-5, L["act"], L["slope"]
    _cabi.check(_cabi.lib().gb_in_bwd(C.byref(p), torch.cuda.current_stream().cuda_stream), "gb_in_bwd")


def fresh(L):
    N, Cc, H, W = L["shape"]
    return dict(dx=torch.empty_like(L["x"]), dsum=torch.ones(N, 1, H, W, Cc, device=dev) if L["res"] else None,
                bstats=torch.zeros(N * Cc * 2 + 4, device=dev), dbias=torch.zeros(Cc, device=dev))


def time_us(L, cold, reps=20, warm=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = fresh(L)
    ts = []
    for i in range(warm + reps):
        out["bstats"].zero_()
        if cold:
            flush.zero_()
        else:
            L["dy"].mul_(1.0)  # the producer just wrote the gradient: it sits in L2 as far as it fits
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch(L, out)
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    lib = _cabi.lib()
    layers = [
        make("G c7 / u2   64ch 256x256 relu", B, 64, 256, 256, ACT_RELU, 0, False),
        make("G u2->c7out 64ch 256x256 relu border3", B, 64, 256, 256, ACT_RELU, 3, False),
        make("G d1 / u1  128ch 128x128 relu", B, 128, 128, 128, ACT_RELU, 0, False),
        make("G resblock 256ch 64x64 relu border1", B, 256, 64, 64, ACT_RELU, 1, False),
        make("G resblock 256ch 64x64 none border1 +res", B, 256, 64, 64, ACT_NONE, 1, True),
        make("D l2       128ch 64x64 leaky", B, 128, 64, 64, ACT_LEAKY, 0, False),
        make("D l3       256ch 32x32 leaky", B, 256, 32, 32, ACT_LEAKY, 0, False),
        make("D l4       512ch 31x31 leaky", B, 512, 31, 31, ACT_LEAKY, 0, False),
    ]
    print(f"device: {torch.cuda.get_device_name(0)}  batch {B}")
    print(f"{'layer':44s} {'variant':18s} {'cold us':>9s} {'GB/s':>8s} {'hot us':>9s} {'GB/s':>8s}  {'max|ddx|':>9s} served-by")
    for L in layers:
        ref = None
        for vname, knobs in VARIANTS:
            for k in (22, 6, 24):
                lib.gb_debug_knob(k, 0)
            for k, v in knobs.items():
                lib.gb_debug_knob(k, v)
            lib.gb_debug_knob(23, 0)
            lib.gb_debug_knob(25, 0)
            out = fresh(L)
            launch(L, out)
            torch.cuda.synchronize()
            served = f"gen2={lib.gb_debug_knob(23, 0)} onchip={lib.gb_debug_knob(25, 0)}"
            got = (out["dx"].float(), out["dbias"].clone(), out["dsum"].clone() if L["res"] else None)
            if ref is None:
                ref, err = got, 0.0
            else:
                err = (got[0] - ref[0]).abs().max().item() / max(ref[0].abs().max().item(), 1e-12)
                ok = err <= 2.0 ** -7 and torch.allclose(got[1], ref[1], rtol=2e-3, atol=2e-3 * ref[1].abs().max().item())
                if L["res"]:
                    ok = ok and torch.allclose(got[2], ref[2], rtol=1e-5, atol=1e-5)
                if not ok:
                    print(f"MISMATCH {L['name']} {vname}: rel dx err {err:.3e}")
            tc, th = time_us(L, True), time_us(L, False)
            print(f"{L['name']:44s} {vname:18s} {tc:9.1f} {L['bytes'] / tc / 1e3:8.0f} {th:9.1f} {L['bytes'] / th / 1e3:8.0f}  {err:9.2e} {served}")
        for k in (22, 6, 24):
            lib.gb_debug_knob(k, 0)


def fwd_table(B):
    """Forward: first-generation fast kernel against the second generation (knob 26); 4 B per element algorithmic
    (+ 2 B residual, + the border of the result)."""
    lib = _cabi.lib()
    shapes = [("G c7 / u2   64ch 256x256 relu", 64, 256, 256, ACT_RELU, 0, False), ("G u2->out   64ch 256x256 relu border3", 64, 256, 256, ACT_RELU, 3, False),
              ("G d1 / u1  128ch 128x128 relu", 128, 128, 128, ACT_RELU, 0, False), ("G resblock 256ch 64x64 relu border1", 256, 64, 64, ACT_RELU, 1, False),
              ("G resblock 256ch 64x64 none border1 +res", 256, 64, 64, ACT_NONE, 1, True), ("D l2       128ch 64x64 leaky", 128, 64, 64, ACT_LEAKY, 0, False)]
    print(f"\n{'forward layer':44s} {'variant':10s} {'cold us':>9s} {'GB/s':>8s} {'hot us':>9s} {'GB/s':>8s}  {'max|dy|':>9s} served-by-gen2")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for name, Cc, H, W, act, yp, res in shapes:
        x = (torch.randn(B, 1, H, W, Cc, device=dev) * 1.3 + 0.4).to(torch.bfloat16)
        xf = x.float()
        stats = torch.stack([xf.sum(dim=(1, 2, 3)), (xf * xf).sum(dim=(1, 2, 3))], dim=-1).contiguous()
        r = torch.randn(B, 1, H, W, Cc, device=dev).to(torch.bfloat16) if res else None
        y = torch.empty(B, 1, H + 2 * yp, W + 2 * yp, Cc, device=dev, dtype=torch.bfloat16)
        nbytes = x.numel() * 2 * (3 if res else 2)

        def launch():
            p = _cabi.InFwdParams()
            p.x, p.y = ops.make_view(x), ops.make_view(y, yp)
            if res:
                p.res = ops.make_view(r)
            p.stats, p.eps, p.act, p.act_slope = stats.data_ptr(), 1e-5, act, 0.2 if act == ACT_LEAKY else 0.0
            _cabi.check(lib.gb_in_fwd(C.byref(p), torch.cuda.current_stream().cuda_stream), "gb_in_fwd")

        ref = None
        for vname, k26 in (("gen1", 0), ("gen2", 1)):
            lib.gb_debug_knob(26, k26)
            lib.gb_debug_knob(27, 0)
            y.fill_(float("nan"))
            launch()
            torch.cuda.synchronize()
            served = lib.gb_debug_knob(27, 0)
            got = y.float().clone()
            err = 0.0 if ref is None else (got - ref).abs().max().item() / ref.abs().max().item()
            if ref is None:
                ref = got
            elif not (err <= 2.0 ** -7):
                print(f"MISMATCH forward {name} {vname}: {err:.3e}")
            ts = {}
            for cold in (True, False):
                t = []
                for i in range(23):
                    if cold:
                        flush.zero_()
                    else:
                        x.mul_(1.0)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    launch()
                    e1.record()
                    torch.cuda.synchronize()
                    if i >= 3:
                        t.append(e0.elapsed_time(e1) * 1e3)
                t.sort()
                ts[cold] = t[len(t) // 2]
            print(f"{name:44s} {vname:10s} {ts[True]:9.1f} {nbytes / ts[True] / 1e3:8.0f} {ts[False]:9.1f} {nbytes / ts[False] / 1e3:8.0f}  {err:9.2e} {served}")
        lib.gb_debug_knob(26, 0)


if __name__ == "__main__":
    main()
    fwd_table(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
