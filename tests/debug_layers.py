"""Bring-up aid: layer-by-layer comparison of one generator forward against the bf16-rounding-point oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ganslate_b200.nn import layers
from ganslate_b200.nn.generators import Resnet2D
from ganslate_b200.nn.discriminators import PatchGAN2D
from oracle import torch_oracle as O


def main(kind="G", size=256):
    torch.manual_seed(0)
    if kind == "G":
        ref = O.init_weights(O.OracleResnet2D(3, 3, 9))
        ours = Resnet2D(3, 3, "instance", 9)
    else:
        ref = O.init_weights(O.OraclePatchGAN2D(3))
        ours = PatchGAN2D(3, 64, 3, (4, 4), "instance")
    ours.load_state_dict(ref.state_dict())
    ours = ours.cuda()
    x = torch.rand(1, 3, size, size) * 2 - 1
    rec = []
    sc, sn = layers.step_conv, layers.step_norm_act

    def conv_w(tape, b, m, act=0, slope=0.0):
        out = sc(tape, b, m, act, slope)
        rec.append(("raw" if out.raw else "act", out))
        return out

    def norm_w(*a, **k):
        out = sn(*a, **k)
        rec.append(("act", out))
        return out

    layers.step_conv, layers.step_norm_act = conv_w, norm_w
    y = ours(x.cuda().requires_grad_(True))
    layers.step_conv, layers.step_norm_act = sc, sn
    O.TRACE = []
    yr = O.forward_bf16_points(ref, x)
    tr = O.TRACE
    O.TRACE = None
    print("groups ours", len(rec), "oracle", len(tr))
    for i, ((k1, b), (k2, t)) in enumerate(zip(rec, tr)):
        p = b.pad
        tt = b.t[:, 0, p:b.t.shape[2] - p, p:b.t.shape[3] - p, :b.channels].permute(0, 3, 1, 2).float().cpu()
        d = (tt - t)
        nz = (d.abs() > 0).float().mean().item()
        print(f"{i:3d} {k1:4s}/{k2:4s} shape {tuple(t.shape)} rel_l2 {(d.norm() / t.norm()).item():.3e} "
              f"max {d.abs().max().item():.3e} frac_diff {nz:.4f} ref_rms {t.pow(2).mean().sqrt().item():.3e}")
    print("final", ((y.detach().cpu() - yr).norm() / yr.norm()).item())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "G", int(sys.argv[2]) if len(sys.argv) > 2 else 256)
