import sys, os
sys.path.insert(0, "/root/repo")
import torch
from ganslate_b200 import ops, _cabi
from tools.conv_microbench import time_us_graph
dev = "cuda"
lib = _cabi.lib()
def run(name, op, xin_shape, wshape, with_stats):
    x = torch.randn(*xin_shape, device=dev).to(torch.bfloat16)
    w = torch.randn(*wshape, device=dev) * 0.05
    b = torch.zeros(op.cout, device=dev)
    stats = torch.zeros(1, op.cout_pad, 2, device=dev) if with_stats else None
    lib.gb_debug_knob(15, 0)
    t = time_us_graph(lambda: op.run_fwd(ops.make_view(x), dev, w, b, stats=stats))
    print(f"{name:46s} stats={with_stats!s:5s} {t:8.1f} us [k{lib.gb_debug_knob(15, 0)}]", flush=True)
for st in (True, False):
    op = ops.ConvOp(64, 16, (2, 2, 2), (2, 2, 2), (0, 0, 0), transposed=True)
    run("up convT k2 s2 64->16, 16x128x128 -> 32x256x256", op, (1, 16, 128, 128, 64), (64, 16, 2, 2, 2), st)
    op = ops.ConvOp(16, 32, (2, 2, 2), (2, 2, 2), (0, 0, 0))
    run("down conv k2 s2 16->32, 32x256x256 -> 16x128x128", op, (1, 32, 256, 256, 16), (32, 16, 2, 2, 2), st)
    op = ops.ConvOp(32, 1, (1, 1, 1), (1, 1, 1), (0, 0, 0))
    run("out conv 1x1 32->1, 32x256x256", op, (1, 32, 256, 256, 32), (1, 32, 1, 1, 1), st)
