"""Times single convolution launches (forward / data gradient / weight gradient) of the layers that dominate a
CycleGAN step, under bring-up knob variants, with CUDA events on the launch stream.  Not a bench value: this is the
A/B tool used to pick kernel heuristics (results are copied into profiles/ by hand).

    python tools/conv_microbench.py [batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ganslate_b200 import _cabi, ops

dev = "cuda"


def time_us(fn, reps=20, warm=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def time_us_graph(fn, n=10, reps=5):
    """Per-launch time of `n` back-to-back launches replayed from a CUDA graph: no host launch latency inside the timed
    region (time_us brackets ONE eager launch with events, and the ~20 us of Python between `e0.record()` and the kernel
    reaching the stream are then part of the number whenever the GPU is idle -- it is, after the L2 flush).  Warm L2:
    what a layer sees inside a step."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / n)
    ts.sort()
    return ts[len(ts) // 2]


def layer(name, cin, cout, k, stride, pad, H, W, N, border=0, transposed=False, op_pad=0):
    """border: materialised reflection border of the input buffer (the conv itself then has padding 0)."""
    op = ops.ConvOp(cin, cout, (1, k, k), (1, stride, stride), (0, pad, pad), transposed=transposed,
                    output_padding=(0, op_pad, op_pad))
    wshape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    w = (torch.randn(wshape, device=dev) * 0.02)
    bias = torch.zeros(cout, device=dev)
    x = torch.randn(N, 1, H + 2 * border, W + 2 * border, op.cin_pad, device=dev).to(torch.bfloat16)
    xv = ops.make_view(x)
    od, oh, ow = op.out_extent((1, H + 2 * border, W + 2 * border))
    dy = torch.randn(N, od, oh, ow, op.cout_pad, device=dev).to(torch.bfloat16)
    dx = torch.zeros(N, 1, H + 2 * border, W + 2 * border, op.cin_pad, device=dev, dtype=torch.float32)
    stats = torch.zeros(N, op.cout_pad, 2, device=dev)
    fl = op.flops((1, H + 2 * border, W + 2 * border), N)
    # gradient-side pixel windows read dOut from a zero-bordered copy (the pad kernel is part of the real step, not of
    # this launch timing)
    dyv = torch.nn.functional.pad(dy, (0, 0, ops.BWD_BORDER, ops.BWD_BORDER)) if op.bwd_window else ops.make_view(dy)
    return dict(name=name, op=op, w=w, bias=bias, xv=xv, x=x, dy=dy, dyv=dyv, dxv=ops.make_view(dx), dx=dx,
                stats=stats, flops=fl)


NO_STATS = False


def run(L, what):
    op = L["op"]
    if what == "fwd":
        return lambda: op.run_fwd(L["xv"], dev, L["w"], L["bias"], stats=None if NO_STATS else L["stats"])
    if what == "dgrad":
        return lambda: op.run_dgrad(L["dyv"], L["w"], L["dxv"], accumulate=False)
    return lambda: op.run_wgrad(L["xv"], L["dyv"], L["w"].shape, dev)


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("batch", type=int, nargs="?", default=8)
    ap.add_argument("--layers", default="", help="comma-separated layer indices (default: all)")
    ap.add_argument("--variants", default="", help="comma-separated variant indices (default: all)")
    ap.add_argument("--what", default="fwd,dgrad,wgrad")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-stats", action="store_true", help="forward without the InstanceNorm statistics epilogue")
    ap.add_argument("--graph", action="store_true", help="time 10 back-to-back launches replayed from a CUDA graph (no host "
                                                         "launch latency in the number; warm L2)")
    args = ap.parse_args()
    B = args.batch
    global NO_STATS
    NO_STATS = args.no_stats
    lib = _cabi.lib()
    layers = [
        layer("res 3x3 256->256 64x64 (+border 1)", 256, 256, 3, 1, 0, 64, 64, B, border=1),
        layer("D 4x4 s1 256->512 32->31", 256, 512, 4, 1, 1, 32, 32, B),
        layer("down 3x3 s2 64->128 256->128", 64, 128, 3, 2, 1, 256, 256, B),
        layer("down 3x3 s2 128->256 128->64", 128, 256, 3, 2, 1, 128, 128, B),
        layer("up convT 3x3 s2 256->128 64->128", 256, 128, 3, 2, 1, 64, 64, B, transposed=True, op_pad=1),
        layer("up convT 3x3 s2 128->64 128->256", 128, 64, 3, 2, 1, 128, 128, B, transposed=True, op_pad=1),
        layer("first 7x7 3->64 256x256 (+border 3)", 3, 64, 7, 1, 0, 256, 256, B, border=3),
        layer("last 7x7 64->3 256x256 (+border 3)", 64, 3, 7, 1, 0, 256, 256, B, border=3),
    ]
    variants = [
        ("per-tap, power-of-two patches", {9: 1, 0: 1}),
        ("per-tap, free patch shape", {9: 1}),
        ("default", {}),
        ("pair bn256", {9: 2, 11: 256}),
        ("pair bn128", {9: 2, 11: 128}),
        ("pair bn256 pitch16", {9: 2, 11: 256, 10: 1}),
        ("wgrad two row tiles", {12: 2}),
        ("cg2 pair-MMA persistent (knob 16)", {16: 1}),
        ("cg2 bn128", {16: 1, 1: 128}),
        ("cg2 3-stage ring", {16: 1, 17: 3}),
        ("persistent 1-CTA (knob 16 = 2)", {16: 2}),
        ("per-tap, per-thread store epilogue", {9: 1, 29: 1}),
        ("per-tap, bulk-store epilogue", {9: 1, 29: 2}),
        ("per-tap, coalesced-store epilogue", {9: 1, 29: 3}),
        ("persistent (knob 16 = 3)", {9: 1, 16: 3}),
        ("persistent, 74 CTAs", {9: 1, 16: 3, 18: 74}),
        ("wgrad per-thread red.global.add epilogue", {12: 4}),
    ]
    if args.layers:
        layers = [layers[int(i)] for i in args.layers.split(",")]
    if args.variants:
        variants = [variants[int(i)] for i in args.variants.split(",")]
    print(f"batch {B}; us per launch, TFLOP/s algorithmic; " + ("10 launches back to back from a CUDA graph, warm L2" if args.graph
          else f"median of {args.reps} single eager launches, L2 flushed (includes host launch latency when the GPU idles)"))
    for L in layers:
        print(f"--- {L['name']}  ({L['flops'] / 1e9:.2f} GFLOP)")
        for vname, knobs in variants:
            old = {k: lib.gb_debug_knob(k, v) for k, v in knobs.items()}
            try:
                row = []
                for what in args.what.split(","):
                    lib.gb_debug_knob(15, 0)
                    lib.gb_debug_knob(14, 0)
                    t = time_us_graph(run(L, what)) if args.graph else time_us(run(L, what), reps=args.reps, warm=1 if args.reps < 5 else 3)
                    path = lib.gb_debug_knob(14, 0) if what == "wgrad" else lib.gb_debug_knob(15, 0)
                    row.append(f"{what} {t:7.1f}us {L['flops'] / t / 1e6:6.0f}TF [k{path}]")
                print(f"  {vname:32s} " + "  ".join(row), flush=True)
            finally:
                for k, v in old.items():
                    lib.gb_debug_knob(k, v)


if __name__ == "__main__":
    main()
