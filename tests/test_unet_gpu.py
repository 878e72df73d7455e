"""Unet2D (ganslate/nn/generators/unet/unet2d.py) on the sm_100a kernels against the CPU oracle restatement
(oracle/torch_oracle.py::OracleUnet2D, pinned to the reference module in tests/test_oracle.py), and one Pix2Pix
iteration with the U-Net generator (BASELINE config 2's generator).  Tolerances as for the other networks:
outputs relative L2 <= 3e-2, gradients cosine >= 0.9 / 0.95 against the fp32 oracle, losses <= 1e-2 relative."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _load(ours, ref):
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    assert list(sd.keys()) == list(ours.state_dict().keys())
    ours.load_state_dict(sd)


@pytest.mark.parametrize("num_downs,size,width", [(5, 128, 128), (6, 64, 128)])
def test_unet2d_vs_oracle(num_downs, size, width):
    from ganslate_b200.nn.generators import Unet2D
    from oracle import torch_oracle as O
    from parity_util import cosine, rel_l2
    torch.manual_seed(0)
    ref = O.init_weights(O.OracleUnet2D(3, 3, num_downs, ngf=16))
    ours = Unet2D(3, 3, num_downs, "instance", ngf=16).cuda()
    _load(ours, ref)
    x, _ = O.synthetic_batch(2, 3, size, seed=5, width=width)
    xr, xo = x.clone().requires_grad_(True), x.clone().cuda().requires_grad_(True)
    yr, yo = ref(xr), ours(xo)
    assert yo.shape == yr.shape and rel_l2(yo, yr) <= 3e-2, rel_l2(yo, yr)
    g = torch.randn_like(yr)
    yr.backward(g)
    yo.backward(g.cuda())
    torch.cuda.synchronize()
    assert cosine(xo.grad, xr.grad) >= 0.95
    bad = []
    for (k, p), (_, q) in zip(ref.named_parameters(), ours.named_parameters()):
        if k.endswith("weight"):
            c = cosine(q.grad, p.grad)
            if c < 0.95:
                bad.append((k, c))
    assert not bad, bad


def test_unet2d_dropout_modes():
    """Dropout(0.5) of the intermediate blocks: active in training mode (two calls differ, result finite), absent in
    eval mode (equal to the network without dropout)."""
    from ganslate_b200.nn.generators import Unet2D
    from oracle import torch_oracle as O
    torch.manual_seed(0)
    ref = O.init_weights(O.OracleUnet2D(3, 3, 6, ngf=16, use_dropout=True))
    drop = Unet2D(3, 3, 6, "instance", ngf=16, use_dropout=True).cuda()
    plain = Unet2D(3, 3, 6, "instance", ngf=16, use_dropout=False).cuda()
    _load(drop, ref)
    plain.load_state_dict({k: v.detach().clone() for k, v in ref.state_dict().items()})
    x, _ = O.synthetic_batch(2, 3, 64, seed=6)
    x = x.cuda().requires_grad_(True)
    drop.train()
    y1, y2 = drop(x), drop(x)
    y1.sum().backward()
    torch.cuda.synchronize()
    assert torch.isfinite(y1).all() and torch.isfinite(x.grad).all()
    assert (y1 - y2).abs().max().item() > 1e-4
    drop.eval(), plain.eval()
    with torch.no_grad():
        e1, e2, e3 = drop(x), drop(x), plain(x)
    # (not bit-equal: the InstanceNorm statistics are fp32 atomics, their order can move a bf16 rounding)
    from parity_util import rel_l2
    assert rel_l2(e1, e2) < 5e-3 and rel_l2(e1, e3) < 5e-3


def test_pix2pix_unet_step_vs_oracle():
    from ganslate_b200.presets import pix2pix_unet2d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle as O
    from parity_util import cosine, rel_l2
    oracle = O.OraclePix2Pix(lambda_pix2pix=30.0, n_layers=3, seed=0, unet=dict(num_downs=5, ngf=16))
    torch.manual_seed(0)
    ours = build_gan(pix2pix_unet2d(batch_size=2, num_downs=5, ngf=16, use_dropout=False, n_layers=3))
    for name in ("G", "D"):
        for (k1, p1), (k2, p2) in zip(oracle.networks[name].state_dict().items(), ours.networks[name].state_dict().items()):
            assert k1 == k2 and torch.equal(p1, p2.cpu()), (name, k1)
    a, b = O.synthetic_batch(2, 3, 128, seed=1, width=64)
    lo = oracle.optimize_parameters(a, b, step_optimizers=False)
    for o in ours.optimizers.values():
        o.step = lambda *a, **k: None
    ours.set_input({"A": a, "B": b})
    ours.optimize_parameters()
    torch.cuda.synchronize()
    for k, v in lo.items():
        assert abs(float(ours.losses[k]) - v) <= 1e-2 * abs(v), (k, v, float(ours.losses[k]))
    assert rel_l2(ours.visuals["fake_B"], oracle.visuals["fake_B"]) < 3e-2
    for name in ("G", "D"):
        po, pg = dict(oracle.networks[name].named_parameters()), dict(ours.networks[name].named_parameters())
        for k in po:
            if k.endswith("weight"):
                assert cosine(pg[k].grad, po[k].grad) > 0.9, (name, k)


def test_unet3d_vs_oracle():
    """Unet3D (ganslate/nn/generators/unet/unet3d.py): the same block over 3-D layers (k4 s2 p1 convolutions and
    transposed convolutions = 8 parity classes of 8 taps, InstanceNorm3d, channel-slice concatenation)."""
    from ganslate_b200.nn.generators import Unet3D
    from oracle import torch_oracle as O
    from parity_util import cosine, rel_l2
    torch.manual_seed(0)
    ref = O.init_weights(O.OracleUnet3D(1, 1, 5, ngf=8))
    ours = Unet3D(1, 1, 5, "instance", ngf=8).cuda()
    _load(ours, ref)
    g0 = torch.Generator().manual_seed(9)
    x = torch.rand((1, 1, 32, 32, 64), generator=g0) * 2 - 1
    xr, xo = x.clone().requires_grad_(True), x.clone().cuda().requires_grad_(True)
    yr, yo = ref(xr), ours(xo)
    assert yo.shape == yr.shape and rel_l2(yo, yr) <= 3e-2, rel_l2(yo, yr)
    g = torch.randn_like(yr)
    yr.backward(g)
    yo.backward(g.cuda())
    torch.cuda.synchronize()
    assert cosine(xo.grad, xr.grad) >= 0.95
    bad = []
    for (k, p), (_, q) in zip(ref.named_parameters(), ours.named_parameters()):
        if k.endswith("weight"):
            c = cosine(q.grad, p.grad)
            if c < 0.95:
                bad.append((k, c))
    assert not bad, bad
