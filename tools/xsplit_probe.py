"""Probe: the 7x7 64 -> 3 output layer re-described as a (7 x 1) convolution to 21 pseudo-channels (dx, co) + a shift-add
over dx.  Times stage 1 on the existing kernels (fp32 and bf16 destinations) from a CUDA graph."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ganslate_b200 import _cabi, ops
import conv_microbench as mb

dev = "cuda"
B = 8
lib = _cabi.lib()
x = torch.randn(B, 1, 262, 262, 64, device=dev).to(torch.bfloat16)
for cout, kern, pad, name in ((21, (1, 7, 1), (0, 0, 0), "stage 1 of the forward (rows y..y+6 of the bordered input)"),
                              (21, (1, 7, 1), (0, 3, 0), "stage 1 of the first layer's data gradient (zero padding in y)"),
                              (3, (1, 7, 7), (0, 0, 0), "the layer as it runs now")):
    op = ops.ConvOp(64, cout, kern, (1, 1, 1), pad)
    w = torch.randn((cout, 64) + kern, device=dev) * 0.02
    od, oh, ow = op.out_extent((1, 262, 262))
    for fp32 in (True, False):
        y = torch.empty((B, od, oh, ow, op.cout_pad), dtype=torch.float32 if fp32 else torch.bfloat16, device=dev)
        fn = lambda: op.run_fwd_into(ops.make_view(x), w, None, ops.make_view(y), out_fp32=fp32)
        lib.gb_debug_knob(15, 0)
        t = mb.time_us_graph(fn)
        print(f"{name:70s} cout {cout:2d} out {'fp32' if fp32 else 'bf16'} {tuple(y.shape)}: {t:7.1f} us  [k{lib.gb_debug_knob(15, 0)}]", flush=True)
