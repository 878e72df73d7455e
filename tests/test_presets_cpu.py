"""Every bench.py workload preset (BASELINE.json configurations 2, 4, 5 -- the others have their own recipe tests in
tests/test_host_networks_cpu.py) builds through the plug-in builders and completes one iteration of host logic on the
CPU restatement of the ABI at a small size: a wrong config key, channel plumbing or tape construction shows up here,
not in the GPU visit that measures the workload."""
import math
import random

import pytest
import torch

import fake_cabi

CASES = [
    ("pix2pix_unet2d", dict(batch_size=1, num_downs=5, ngf=8, n_layers=3), (3, 64, 96), {"G", "D", "pix2pix"}),
    ("cyclegan_vnet3d", dict(first_layer_channels=8, ndf=8, n_layers=2), (1, 16, 32, 32), {"G_AB", "G_BA", "D_A", "D_B", "cycle_A", "cycle_B"}),
    ("revgan_vnet3d", dict(channels=2, first_layer_channels=8, ndf=8, n_layers=2), (2, 16, 32, 32), {"G_AB", "G_BA", "D_A", "D_B", "cycle_A", "cycle_B"}),
]


@pytest.mark.parametrize("name,kw,shape,expect", CASES, ids=[c[0] for c in CASES])
def test_preset_runs_one_iteration_of_host_logic(monkeypatch, name, kw, shape, expect):
    fake_cabi.install(monkeypatch)
    from ganslate_b200 import presets
    from ganslate_b200.nn.gans import base
    from ganslate_b200.utils.builders import build_gan
    monkeypatch.setattr(base.BaseGAN, "_specify_device", lambda self: torch.device("cpu"))
    torch.manual_seed(0)
    random.seed(0)
    gan = build_gan(getattr(presets, name)(**kw))
    for o in gan.optimizers.values():
        monkeypatch.setattr(o, "step", lambda *a, **k: None)
    a, b = torch.rand((1,) + shape) * 2 - 1, torch.rand((1,) + shape) * 2 - 1
    gan.set_input({"A": a, "B": b})
    gan.optimize_parameters()
    losses = {k: float(v.detach()) for k, v in gan.losses.items() if v is not None}
    assert expect <= set(losses) and all(math.isfinite(v) for v in losses.values()), losses
    for net_name, net in gan.networks.items():
        got = [p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in net.parameters()]
        assert all(got), (net_name, got.count(False))


def test_direct_parameter_gradients_on_a_shared_reversible_generator(monkeypatch):
    """ops.DIRECT_PARAM_GRAD on RevGAN + Vnet3D: ONE generator serves both directions (forward and inverse) and is used
    four times per backward, with retain_graph between the generator and discriminator passes -- the kernels'
    accumulation into `param.grad` must leave exactly the gradients autograd's accumulation leaves."""
    fake_cabi.install(monkeypatch)
    from ganslate_b200 import ops, presets
    from ganslate_b200.nn.gans import base
    from ganslate_b200.utils.builders import build_gan
    monkeypatch.setattr(base.BaseGAN, "_specify_device", lambda self: torch.device("cpu"))
    grads = {}
    for direct in (False, True):
        monkeypatch.setattr(ops, "DIRECT_PARAM_GRAD", direct)
        torch.manual_seed(0)
        random.seed(0)
        gan = build_gan(presets.revgan_vnet3d(channels=2, first_layer_channels=8, ndf=8, n_layers=2))
        for o in gan.optimizers.values():
            monkeypatch.setattr(o, "step", lambda *a, **k: None)
        g = torch.Generator().manual_seed(3)
        a, b = torch.rand((1, 2, 16, 32, 32), generator=g) * 2 - 1, torch.rand((1, 2, 16, 32, 32), generator=g) * 2 - 1
        gan.set_input({"A": a, "B": b})
        gan.optimize_parameters()
        grads[direct] = {(n, k): p.grad.clone() for n, net in gan.networks.items() for k, p in net.named_parameters()}
    assert grads[False].keys() == grads[True].keys() and len(grads[True]) > 100
    for k, v in grads[False].items():
        assert torch.allclose(grads[True][k], v, rtol=1e-5, atol=1e-7 * max(1.0, float(v.abs().max()))), k
