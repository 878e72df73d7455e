"""Pix2Pix (conditional GAN) iteration vs the CPU oracle; tolerances as in test_cyclegan_gpu.py."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_pix2pix_step_vs_oracle():
    from ganslate_b200.presets import pix2pix_resnet2d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle as O
    from parity_util import cosine, rel_l2
    oracle = O.OraclePix2Pix(lambda_pix2pix=30.0, n_residual_blocks=9, n_layers=3, seed=0)
    torch.manual_seed(0)
    ours = build_gan(pix2pix_resnet2d(batch_size=2, n_layers=3))
    for name in ("G", "D"):
        for (k1, p1), (k2, p2) in zip(oracle.networks[name].state_dict().items(), ours.networks[name].state_dict().items()):
            assert k1 == k2 and torch.equal(p1, p2.cpu()), (name, k1)
    a, b = O.synthetic_batch(2, 3, 128, seed=1, width=64)   # cityscapes-like 2:1 aspect, scaled down
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
