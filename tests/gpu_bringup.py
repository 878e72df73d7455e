"""Bring-up script (not a pytest file): runs every kernel family against torch fp32 on the GPU and prints an
error table without stopping at the first failure.  Usage: python tests/gpu_bringup.py"""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from ganslate_b200 import ops, _cabi
from ganslate_b200.nn import layers

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"


def bf(x):
    return x.to(torch.bfloat16).float()


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def conv_case(name, cin, cout, k, s, p, H, W, N=1, transposed=False, op_pad=0, reflect=0, bias=True, D=None, act=None):
    torch.manual_seed(0)
    dims = 3 if D else 2
    if transposed:
        mod = (layers.ConvTranspose3d if D else layers.ConvTranspose2d)(cin, cout, k, stride=s, padding=p,
                                                                        output_padding=op_pad, bias=bias).to(dev)
    else:
        mod = (layers.Conv3d if D else layers.Conv2d)(cin, cout, k, stride=s, padding=p, bias=bias).to(dev)
    with torch.no_grad():
        mod.weight.copy_(bf(torch.randn_like(mod.weight) * 0.1))
        if bias:
            mod.bias.copy_(torch.randn_like(mod.bias) * 0.1)
    shape = (N, cin, D, H, W) if D else (N, cin, H, W)
    x = bf(torch.randn(shape, device=dev)).requires_grad_(True)
    # ours
    mods = ([layers.ReflectionPad2d(reflect)] if reflect else []) + [mod] + ([layers.Tanh()] if act == "tanh" else [])
    y = _Holder(mods).to(dev)(x)
    # reference
    xr = x.detach().clone().requires_grad_(True)
    wr = mod.weight.detach().clone().requires_grad_(True)
    br = mod.bias.detach().clone().requires_grad_(True) if bias else None
    xin = F.pad(xr, (reflect,) * 4, mode="reflect") if reflect else xr
    if transposed:
        fn = F.conv_transpose3d if D else F.conv_transpose2d
        yr = fn(xin, wr, br, stride=s, padding=p, output_padding=op_pad)
    else:
        fn = F.conv3d if D else F.conv2d
        yr = fn(xin, wr, br, stride=s, padding=p)
    if act == "tanh":
        yr = torch.tanh(yr)
    g = bf(torch.randn_like(yr))
    y.backward(g)
    yr.backward(g)
    torch.cuda.synchronize()
    e_y = rel(y, yr)
    e_dx = rel(x.grad, xr.grad)
    e_dw = rel(mod.weight.grad, wr.grad)
    e_db = rel(mod.bias.grad, br.grad) if bias else 0.0
    ok = e_y < 1e-2 and e_dx < 1e-2 and e_dw < 1e-2 and e_db < 1e-2
    print(f"{'OK  ' if ok else 'FAIL'} {name:34s} y {e_y:.2e} dx {e_dx:.2e} dw {e_dw:.2e} db {e_db:.2e}", flush=True)
    return ok


class _Holder(torch.nn.Module):
    def __init__(self, mods):
        super().__init__()
        self.model = torch.nn.Sequential(*mods)

    def forward(self, x):
        return layers.run_network(self, list(self.model), x)


def seq_case(name, ours_mods, ref_mods, shape, tol=3e-2):
    """ours_mods / ref_mods: parallel module lists (ganslate_b200 layers vs torch.nn); weights copied ours -> ref."""
    torch.manual_seed(3)
    ours = _Holder(ours_mods).to(dev)
    ref = torch.nn.Sequential(*ref_mods).to(dev)
    with torch.no_grad():
        for p in ours.parameters():
            p.copy_(bf(torch.randn_like(p) * (0.05 if p.dim() > 1 else 0.1)))
    ref.load_state_dict({k.replace("model.", "", 1): v for k, v in ours.state_dict().items()})
    x = bf(torch.randn(shape, device=dev)).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    y = ours(x)
    # reference: the same torch modules walked with round-to-bf16 at this path's storage points, so that ReLU
    # flips caused by the precision choice do not mask kernel errors
    from oracle import torch_oracle as O
    yr = O._seq_bf16(list(ref), O._rb(xr))
    g = torch.randn_like(yr)
    y.backward(g)
    yr.backward(g)
    torch.cuda.synchronize()
    # relative L2 (a single flipped ReLU changes one element by O(max), max-relative would measure that)
    def l2(a, b):
        return ((a - b).norm() / (b.norm() + 1e-20)).item()
    errs = {"y": l2(y, yr), "dx": l2(x.grad, xr.grad)}
    wmax = max(pr.grad.abs().max().item() for pr in ref.parameters() if pr.dim() > 1)
    for (k, p), (_, pr) in zip(ours.named_parameters(), ref.named_parameters()):
        if pr.dim() > 1:
            errs["d" + k.replace("model.", "")] = l2(p.grad, pr.grad)
        else:  # biases (those in front of an InstanceNorm have a mathematically zero gradient): absolute check
            errs["d" + k.replace("model.", "")] = (p.grad - pr.grad).abs().max().item() / wmax
    ok = all(v < tol for v in errs.values())
    print(f"{'OK  ' if ok else 'FAIL'} {name:34s} " + " ".join(f"{k} {v:.1e}" for k, v in errs.items()), flush=True)
    return ok


def in_case(name, C, H, W, N, act, reflect_next):
    L, nn = layers, torch.nn
    a_o = {"relu": [L.ReLU(True)], "leaky": [L.LeakyReLU(0.2, True)], "none": []}[act]
    a_r = {"relu": [nn.ReLU()], "leaky": [nn.LeakyReLU(0.2)], "none": []}[act]
    tail_o = ([L.ReflectionPad2d(reflect_next)] if reflect_next else []) + [L.Conv2d(C, 16, 3, padding=0 if reflect_next else 1)]
    tail_r = ([nn.ReflectionPad2d(reflect_next)] if reflect_next else []) + [nn.Conv2d(C, 16, 3, padding=0 if reflect_next else 1)]
    return seq_case(name, [L.Conv2d(8, C, 3, padding=1), L.InstanceNorm2d(C)] + a_o + tail_o,
                    [nn.Conv2d(8, C, 3, padding=1), nn.InstanceNorm2d(C)] + a_r + tail_r, (N, 8, H, W))


def resblock_case():
    from ganslate_b200.nn.generators.resnet.resnet2d import ResidualBlock
    from oracle.torch_oracle import OracleResBlock
    L, nn = layers, torch.nn
    return seq_case("2 residual blocks C64 24x24", [L.Conv2d(8, 64, 3, padding=1), L.InstanceNorm2d(64), L.ReLU(True),
                                                    ResidualBlock(64, "instance"), ResidualBlock(64, "instance"),
                                                    L.Conv2d(64, 8, 3, padding=1)],
                    [nn.Conv2d(8, 64, 3, padding=1), nn.InstanceNorm2d(64), nn.ReLU(), OracleResBlock(64), OracleResBlock(64),
                     nn.Conv2d(64, 8, 3, padding=1)], (2, 8, 24, 24))


def loss_case():
    torch.manual_seed(2)
    p = torch.randn(2, 1, 30, 30, device=dev, requires_grad=True)
    l = ops.MseConstFn.apply(p, 1.0)
    l.backward()
    pr = p.detach().clone().requires_grad_(True)
    lr = F.mse_loss(pr, torch.ones_like(pr))
    lr.backward()
    a = torch.randn(2, 3, 64, 64, device=dev, requires_grad=True)
    b = torch.randn(2, 3, 64, 64, device=dev)
    l1 = ops.L1Fn.apply(a, b)
    (l1 * 10).backward()
    ar = a.detach().clone().requires_grad_(True)
    l1r = F.l1_loss(ar, b)
    (l1r * 10).backward()
    torch.cuda.synchronize()
    ok = abs(l.item() - lr.item()) < 1e-5 and rel(p.grad, pr.grad) < 1e-5 and abs(l1.item() - l1r.item()) < 1e-5 and rel(
        a.grad, ar.grad) < 1e-5
    print(f"{'OK  ' if ok else 'FAIL'} losses mse {l.item():.6f}/{lr.item():.6f} l1 {l1.item():.6f}/{l1r.item():.6f}")
    return ok


CASES = [
    lambda: conv_case("1x1 64->64 32x32", 64, 64, 1, 1, 0, 32, 32),
    lambda: conv_case("1x1 64->256 32x32", 64, 256, 1, 1, 0, 32, 32),
    lambda: conv_case("1x1 128->128 16x16 (K=128)", 128, 128, 1, 1, 0, 16, 16),
    lambda: conv_case("3x3 s1 p1 64->64 32x32", 64, 64, 3, 1, 1, 32, 32),
    lambda: conv_case("3x3 reflect1 256->256 64x64 (K3)", 256, 256, 3, 1, 0, 64, 64, reflect=1),
    lambda: conv_case("3x3 s2 p1 64->128 64x64 (K2)", 64, 128, 3, 2, 1, 64, 64),
    lambda: conv_case("7x7 reflect3 3->64 64x64 (K1)", 3, 64, 7, 1, 0, 64, 64, reflect=3),
    lambda: conv_case("7x7 reflect3 64->3 tanh 64x64", 64, 3, 7, 1, 0, 64, 64, reflect=3, act="tanh"),
    lambda: conv_case("convT 3x3 s2 p1 op1 128->64 (K4)", 128, 64, 3, 2, 1, 32, 32, transposed=True, op_pad=1),
    lambda: conv_case("4x4 s2 p1 3->64 (K5a)", 3, 64, 4, 2, 1, 64, 64),
    lambda: conv_case("4x4 s1 p1 256->512 31x31 (K5d)", 256, 512, 4, 1, 1, 32, 32),
    lambda: conv_case("4x4 s1 p1 512->1 (K5e)", 512, 1, 4, 1, 1, 31, 31),
    lambda: conv_case("3x3 N=3 ragged 40->72 19x23", 40, 72, 3, 1, 1, 19, 23, N=3),
    lambda: conv_case("3d 3x3x3 p1 16->32 8x16x16", 16, 32, 3, 1, 1, 16, 16, D=8),
    lambda: conv_case("3d 2x2x2 s2 16->32 8x16x16", 16, 32, 2, 2, 0, 16, 16, D=8),
    lambda: conv_case("3d convT 2x2x2 s2 32->16", 32, 16, 2, 2, 0, 8, 8, D=4, transposed=True),
    lambda: conv_case("3d 4x4x4 s2 p1 1->64", 1, 64, 4, 2, 1, 32, 32, D=8),
    lambda: in_case("IN relu C64 32x32", 64, 32, 32, 2, "relu", 0),
    lambda: in_case("IN relu C256 border1", 256, 16, 16, 2, "relu", 1),
    lambda: in_case("IN leaky C128 31x31", 128, 31, 31, 1, "leaky", 0),
    lambda: in_case("IN relu C64 border3", 64, 24, 24, 1, "relu", 3),
    lambda: in_case("IN none C32 20x20", 32, 20, 20, 3, "none", 0),
    resblock_case,
    loss_case,
]

if __name__ == "__main__":
    print("device:", torch.cuda.get_device_name(0), "lib version", _cabi.lib().gb_version())
    sel = [int(a) for a in sys.argv[1:]] or range(len(CASES))
    bad = 0
    for i in sel:
        try:
            if not CASES[i]():
                bad += 1
        except Exception:
            bad += 1
            print(f"EXC  case {i}")
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e:
                print("device error after exception:", e)
                break
    print("failures:", bad)
