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
    y = layers.run_network(mods, x)
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


def norm_case(name, C, H, W, N, act, reflect_out, residual):
    torch.manual_seed(1)
    x = bf(torch.randn(N, C, H, W, device=dev) * 2 + 0.5).requires_grad_(True)
    r = bf(torch.randn(N, C, H, W, device=dev)).requires_grad_(True) if residual else None
    bx = layers.to_buf(x, 0)
    br = layers.to_buf(r, 1) if residual else None
    act_id, slope = {"none": (0, 0.0), "relu": (1, 0.0), "leaky": (2, 0.2)}[act]
    t = ops.NormActFn.apply(bx.t, br.t if br is not None else None, True, act_id, slope, reflect_out,
                            1 if residual else 0, 1e-5)
    # consume through a border-aware export: fold happens in ToChannelsLast backward only; use plain crop + pad check
    full = t.float()  # (N,1,H+2p,W+2p,C)
    xr = x.detach().clone().requires_grad_(True)
    rr = r.detach().clone().requires_grad_(True) if residual else None
    yr = F.instance_norm(xr, eps=1e-5)
    if act == "relu":
        yr = F.relu(yr)
    elif act == "leaky":
        yr = F.leaky_relu(yr, 0.2)
    if residual:
        yr = yr + rr
    yrp = F.pad(yr, (reflect_out,) * 4, mode="reflect") if reflect_out else yr
    ref_full = yrp.permute(0, 2, 3, 1).unsqueeze(1)
    g = bf(torch.randn_like(ref_full))
    e_y = rel(full[..., :C], ref_full)
    t.backward(g.to(torch.bfloat16) if t.shape[-1] == C else F.pad(g, (0, t.shape[-1] - C)).to(torch.bfloat16))
    ref_full.backward(g)
    torch.cuda.synchronize()
    e_dx = rel(x.grad, xr.grad)
    e_dr = rel(r.grad, rr.grad) if residual else 0.0
    ok = e_y < 1e-2 and e_dx < 2e-2 and e_dr < 1e-2
    print(f"{'OK  ' if ok else 'FAIL'} {name:34s} y {e_y:.2e} dx {e_dx:.2e} dres {e_dr:.2e}", flush=True)
    return ok


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
    lambda: norm_case("IN relu C64 32x32", 64, 32, 32, 2, "relu", 0, False),
    lambda: norm_case("IN relu C256 border1", 256, 16, 16, 2, "relu", 1, False),
    lambda: norm_case("IN none C256 residual border1", 256, 16, 16, 2, "none", 1, True),
    lambda: norm_case("IN leaky C128 31x31", 128, 31, 31, 1, "leaky", 0, False),
    lambda: norm_case("IN relu C64 border3", 64, 24, 24, 1, "relu", 3, False),
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
