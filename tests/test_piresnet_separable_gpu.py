"""Piresnet3D, Vnet3D(is_separable=True) and the replicate-padding kernels on the GPU vs the CPU oracle
(oracle/torch_oracle3d.py, pinned to the reference modules and to tests/golden/piresnet3d_separable_small.json).

First B200 run: the driver's round-1 GPU test (all cases passed); ordinary GPU tests since.
Tolerances as tests/test_3d_gpu.py (bf16 storage, fp32 accumulation)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _load(ours, oracle):
    assert list(ours.state_dict().keys()) == list(oracle.state_dict().keys())
    ours.load_state_dict(oracle.state_dict())


def _compare(ref, ours, x, inverse, out_tol=3e-2, cos_tol=0.9):
    from parity_util import cosine, rel_l2
    xr = x.clone().requires_grad_(True)
    xo = x.clone().cuda().requires_grad_(True)
    kw = {} if inverse is None else {"inverse": inverse}
    yr = ref(xr, **kw)
    yo = ours(xo, **kw)
    assert yo.shape == yr.shape and rel_l2(yo, yr) <= out_tol, (inverse, rel_l2(yo, yr))
    g = torch.randn_like(yr)
    ref.zero_grad()
    ours.zero_grad()
    yr.backward(g)
    yo.backward(g.cuda())
    torch.cuda.synchronize()
    assert cosine(xo.grad, xr.grad) >= cos_tol, (inverse, cosine(xo.grad, xr.grad))
    pr, po = dict(ref.named_parameters()), dict(ours.named_parameters())
    bad = []
    for k, p in pr.items():
        if p.grad is None:
            assert po[k].grad is None or float(po[k].grad.abs().max()) == 0.0, k
            continue
        if k.endswith("weight") and p.dim() > 1 and p.grad.abs().max() > 0:
            c = cosine(po[k].grad, p.grad)
            if c < cos_tol:
                bad.append((k, c))
    assert not bad, (inverse, bad)


@pytest.mark.parametrize("pads,shape", [((1, 1, 1), (2, 16, 4, 6, 5)), ((2, 2, 2), (1, 8, 3, 9, 7)), ((0, 3, 1), (1, 24, 1, 8, 8))])
def test_replicate_pad_kernels_vs_torch(pads, shape):
    """gb_replicate_pad_fwd / _bwd against torch.nn.functional.pad(mode='replicate') and its autograd (exact: copies
    and fp32 sums)."""
    from ganslate_b200 import ops
    N, C, D, H, W = shape
    pz, py, px = pads
    torch.manual_seed(0)
    x = torch.randn(N, D, H, W, C, device="cuda").to(torch.bfloat16)
    y = torch.empty(N, D + 2 * pz, H + 2 * py, W + 2 * px, C, device="cuda", dtype=torch.bfloat16)
    ops.replicate_pad_forward(ops.make_view(x), ops.make_view(y), pads)
    xr = x.float().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    yr = torch.nn.functional.pad(xr, (px, px, py, py, pz, pz), mode="replicate")
    assert torch.equal(y.float().permute(0, 4, 1, 2, 3), yr)
    g = torch.randn_like(yr)
    yr.backward(g)
    gcl = g.permute(0, 2, 3, 4, 1).contiguous()
    dx = torch.full((N, D, H, W, C), 0.5, device="cuda")  # accumulated into
    ops.replicate_pad_backward(ops.make_view(gcl), ops.make_view(dx), pads)
    torch.cuda.synchronize()
    assert torch.allclose(dx.permute(0, 4, 1, 2, 3) - 0.5, xr.grad, rtol=1e-5, atol=1e-5)


def test_piresnet3d_both_directions_vs_oracle():
    from ganslate_b200.nn.generators import Piresnet3D
    from oracle import torch_oracle as O
    from oracle import torch_oracle3d as O3
    torch.manual_seed(0)
    ref = O.init_weights(O3.OraclePiresnet3D(2, 2, 2, first_layer_channels=16, use_inverse=True))
    ours = Piresnet3D(2, 2, "instance", depth=2, first_layer_channels=16, use_memory_saving=False, use_inverse=True).cuda()
    _load(ours, ref)
    x, _ = O3.synthetic_volume(1, 2, 8, 16, seed=3)
    for inverse in (False, True):
        _compare(ref, ours, x, inverse)


def test_separable_vnet3d_both_directions_vs_oracle():
    from ganslate_b200.nn.generators import Vnet3D
    from oracle import torch_oracle as O
    from oracle import torch_oracle3d as O3
    small = dict(first_layer_channels=8, down_blocks=(1, 1), up_blocks=(1, 1))
    torch.manual_seed(0)
    ref = O.init_weights(O3.OracleVnet3D(1, 1, use_inverse=True, is_separable=True, **small))
    ours = Vnet3D(1, 1, "instance", use_memory_saving=False, use_inverse=True, is_separable=True, **small).cuda()
    _load(ours, ref)
    x, _ = O3.synthetic_volume(1, 1, 16, 32, seed=3)
    for inverse in (False, True):
        _compare(ref, ours, x, inverse)


def test_resnet3d_slab_convolutions_vs_oracle():
    """Resnet3D: 7x7x7 layers as seven depth slabs accumulated in FP32 (ops.SlabConv), replicate padding."""
    from ganslate_b200.nn.generators import Resnet3D
    from oracle import torch_oracle as O
    from oracle import torch_oracle3d as O3
    torch.manual_seed(0)
    ref = O.init_weights(O3.OracleResnet3D(1, 2, n_residual_blocks=2))
    ours = Resnet3D(1, 2, "instance", n_residual_blocks=2).cuda()
    _load(ours, ref)
    x, _ = O3.synthetic_volume(1, 1, 8, 16, seed=3)
    _compare(ref, ours, x, None)


def test_revgan_piresnet3d_iteration_vs_oracle():
    """The shipped BraTS experiment's pairing (RevGAN + Piresnet3D + PatchGAN3D(n_layers 2), revgan.yaml:25-39)."""
    import random
    from ganslate_b200.presets import revgan_piresnet3d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle3d as O3
    from parity_util import cosine, rel_l2
    random.seed(0)
    oracle = O3.OracleRevGAN(O3.default_3d_conf(in_channels=1, out_channels=1, ndf=16, n_layers=2, first_layer_channels=16,
                                                piresnet_depth=2), seed=0)
    torch.manual_seed(0)
    ours = build_gan(revgan_piresnet3d(channels=1, depth=2, first_layer_channels=16, ndf=16, n_layers=2))
    for name in ("G", "D_B", "D_A"):
        _load(ours.networks[name], oracle.networks[name])
    a, b = O3.synthetic_volume(1, 1, 16, 32, seed=1)
    lo = oracle.optimize_parameters(a, b, step_optimizers=False)
    for o in ours.optimizers.values():
        o.step = lambda *a, **k: None
    ours.set_input({"A": a, "B": b})
    ours.optimize_parameters()
    torch.cuda.synchronize()
    for k, v in lo.items():
        assert abs(float(ours.losses[k].detach()) - v) <= 2e-2 * abs(v) + 1e-4, (k, v, float(ours.losses[k].detach()))
    for k, tol in (("fake_B", 3e-2), ("fake_A", 3e-2), ("rec_A", 1.2e-1), ("rec_B", 1.2e-1)):
        assert rel_l2(ours.visuals[k], oracle.visuals[k]) <= tol, (k, rel_l2(ours.visuals[k], oracle.visuals[k]))
    bad = []
    for name in ("G", "D_B", "D_A"):
        po, pg = dict(oracle.networks[name].named_parameters()), dict(ours.networks[name].named_parameters())
        for k, p in po.items():
            if k.endswith("weight") and p.dim() > 1 and p.grad is not None and p.grad.abs().max() > 0:
                c = cosine(pg[k].grad, p.grad)
                if c < 0.9:
                    bad.append((name, k, c))
    assert not bad, bad
