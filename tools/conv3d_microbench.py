"""Times single 3-D convolution launches (forward / data gradient / weight gradient) of the V-Net layers that dominate the
3-D CycleGAN / RevGAN steps (BASELINE configs 4 and 5), 10 launches back to back replayed from a CUDA graph.  Not a
bench value: the A/B tool for the 16 / 32-channel k5^3 layers.

    python tools/conv3d_microbench.py [--layers 0,1] [--what fwd,dgrad,wgrad] [--knobs 4=2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ganslate_b200 import _cabi, ops
from tools.conv_microbench import time_us_graph

dev = "cuda"


def layer3d(name, cin, cout, k, ext, N=1):
    op = ops.ConvOp(cin, cout, (k, k, k), (1, 1, 1), (k // 2, k // 2, k // 2))
    w = torch.randn(cout, cin, k, k, k, device=dev) * 0.02
    bias = torch.zeros(cout, device=dev)
    D, H, W = ext
    x = torch.randn(N, D, H, W, op.cin_pad, device=dev).to(torch.bfloat16)
    dy = torch.randn(N, D, H, W, op.cout_pad, device=dev).to(torch.bfloat16)
    dx = torch.zeros(N, D, H, W, ops.pad8(cin), device=dev, dtype=torch.float32)  # (a widened input keeps its 8-channel gradient)
    stats = torch.zeros(N, op.cout_pad, 2, device=dev)
    return dict(name=name, op=op, w=w, bias=bias, xv=ops.make_view(x), dyv=ops.make_view(dy), dxv=ops.make_view(dx),
                stats=stats, flops=op.flops(ext, N), keep=(x, dy, dx))


def run(L, what):
    op = L["op"]
    if what == "fwd":
        return lambda: op.run_fwd(L["xv"], dev, L["w"], L["bias"], stats=L["stats"])
    if what == "dgrad":
        return lambda: op.run_dgrad(L["dyv"], L["w"], L["dxv"], accumulate=False)
    return lambda: op.run_wgrad(L["xv"], L["dyv"], L["w"].shape, dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", default="")
    ap.add_argument("--what", default="fwd,dgrad,wgrad")
    ap.add_argument("--knobs", default="", help="comma-separated knob=value pairs applied to every launch")
    args = ap.parse_args()
    lib = _cabi.lib()
    for kv in filter(None, args.knobs.split(",")):
        k, v = kv.split("=")
        lib.gb_debug_knob(int(k), int(v))
    specs = [
        ("out block k5 32->32 32x256x256", 32, 32, 5, (32, 256, 256)),
        ("up3 coupling k5 16->16 32x256x256", 16, 16, 5, (32, 256, 256)),
        ("in block k5 1->16 32x256x256", 1, 16, 5, (32, 256, 256)),
        ("up2 coupling k5 32->32 16x128x128", 32, 32, 5, (16, 128, 128)),
        ("down0 coupling k5 16->16 16x128x128", 16, 16, 5, (16, 128, 128)),
        ("up1 coupling k5 64->64 8x64x64", 64, 64, 5, (8, 64, 64)),
        ("up0 coupling k5 128->128 4x32x32", 128, 128, 5, (4, 32, 32)),
        ("revgan out block k5 32->32 128^3", 32, 32, 5, (128, 128, 128)),
    ]
    if args.layers:
        specs = [specs[int(i)] for i in args.layers.split(",")]
    print("us per launch (10 launches back to back from a CUDA graph, warm L2), algorithmic TFLOP/s, [kernel path]")
    for s in specs:
        L = layer3d(*s)
        row = []
        for what in args.what.split(","):
            lib.gb_debug_knob(15, 0)
            lib.gb_debug_knob(14, 0)
            t = time_us_graph(run(L, what), n=5 if L["flops"] > 1e11 else 10)
            path = lib.gb_debug_knob(14, 0) if what == "wgrad" else lib.gb_debug_knob(15, 0)
            row.append(f"{what} {t:8.1f}us {L['flops'] / t / 1e6:6.0f}TF [k{path}]")
        print(f"  {L['name']:38s} ({L['flops'] / 1e9:6.1f} GFLOP)  " + "  ".join(row), flush=True)
        del L
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
