"""Times CUT's FeaturePatchMLP launches (csrc/patch_mlp.cu) at the five feature shapes of cut_resnet2d, 10 calls back to
back replayed from a CUDA graph (no host latency in the number).  Not a bench value.

    python tools/patch_mlp_microbench.py [batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ganslate_b200 import ops
from tools.conv_microbench import time_us_graph


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    dev = "cuda"
    print(f"batch {N}; us per call (forward = 1 launch, backward = 3), CUDA-graph replay")
    for C, S in ((3, 262), (128, 256), (256, 128), (256, 64), (256, 64)):
        feat = torch.randn(N, C, S, S, device=dev, requires_grad=True)
        ids = torch.randperm(S * S, device=dev)[:256]
        w1 = (torch.randn(256, C, device=dev) * 0.1).requires_grad_(True)
        b1 = torch.zeros(256, device=dev, requires_grad=True)
        w2 = (torch.randn(256, 256, device=dev) * 0.05).requires_grad_(True)
        b2 = torch.zeros(256, device=dev, requires_grad=True)
        dy = torch.randn(N * 256, 256, device=dev)
        with torch.no_grad():
            tf = time_us_graph(lambda: ops.PatchMlpFn.apply(feat.detach(), ids, w1.detach(), b1.detach(), w2.detach(), b2.detach()))

        def fb():
            y = ops.PatchMlpFn.apply(feat, ids, w1, b1, w2, b2)
            torch.autograd.grad(y, (feat, w1, b1, w2, b2), dy)

        tfb = time_us_graph(fb)
        print(f"  C={C:4d} map {S}x{S}: forward {tf:7.1f} us   forward+backward {tfb:7.1f} us (incl. the zero-fill of dfeat)")


if __name__ == "__main__":
    main()
