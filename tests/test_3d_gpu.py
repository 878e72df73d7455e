"""3-D path on the GPU (Vnet3D both directions, PatchGAN3D, one RevGAN iteration) vs the CPU oracle
(oracle/torch_oracle3d.py, pinned to the reference's modules by tests/test_oracle.py).

Tolerances (bf16 storage, fp32 accumulation): outputs relative L2 <= 3e-2 after one network,
<= 1.2e-1 after two; losses <= 2e-2 relative; weight gradients cosine >= 0.9 against the fp32 oracle."""
import os
import random
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

SMALL = dict(first_layer_channels=8, down_blocks=(1, 1), up_blocks=(1, 1))


def _load(ours, oracle):
    assert list(ours.state_dict().keys()) == list(oracle.state_dict().keys())
    ours.load_state_dict(oracle.state_dict())


def test_vnet3d_both_directions_vs_oracle():
    from ganslate_b200.nn.generators import Vnet3D
    from oracle import torch_oracle as O
    from oracle import torch_oracle3d as O3
    from parity_util import cosine, rel_l2
    torch.manual_seed(0)
    ref = O.init_weights(O3.OracleVnet3D(1, 1, use_inverse=True, **SMALL))
    ours = Vnet3D(1, 1, "instance", use_memory_saving=False, use_inverse=True, **SMALL).cuda()
    _load(ours, ref)
    x, _ = O3.synthetic_volume(1, 1, 16, 32, seed=3)
    for inverse in (False, True):
        xr = x.clone().requires_grad_(True)
        xo = x.clone().cuda().requires_grad_(True)
        yr = ref(xr, inverse=inverse)
        yo = ours(xo, inverse=inverse)
        assert rel_l2(yo, yr) <= 3e-2, (inverse, rel_l2(yo, yr))
        g = torch.randn_like(yr)
        ref.zero_grad()
        ours.zero_grad()
        yr.backward(g)
        yo.backward(g.cuda())
        torch.cuda.synchronize()
        assert cosine(xo.grad, xr.grad) >= 0.9, (inverse, cosine(xo.grad, xr.grad))
        pr, po = dict(ref.named_parameters()), dict(ours.named_parameters())
        bad = []
        for k, p in pr.items():
            if p.grad is None:
                assert po[k].grad is None or float(po[k].grad.abs().max()) == 0.0, k
                continue
            if k.endswith("weight") and p.grad.abs().max() > 0:
                c = cosine(po[k].grad, p.grad)
                if c < 0.9:
                    bad.append((k, c))
        assert not bad, (inverse, bad)


def test_patchgan3d_vs_oracle():
    from ganslate_b200.nn.discriminators import PatchGAN3D
    from oracle import torch_oracle as O
    from oracle import torch_oracle3d as O3
    from parity_util import cosine, rel_l2
    torch.manual_seed(0)
    ref = O.init_weights(O3.OraclePatchGAN3D(1, 16, 2, (4, 4, 4)))
    ours = PatchGAN3D(1, 16, 2, (4, 4, 4), "instance").cuda()
    _load(ours, ref)
    x, _ = O3.synthetic_volume(2, 1, 16, 32, seed=5)
    xr, xo = x.clone().requires_grad_(True), x.clone().cuda().requires_grad_(True)
    yr, yo = ref(xr), ours(xo)
    assert yo.shape == yr.shape and rel_l2(yo, yr) <= 3e-2
    g = torch.randn_like(yr)
    yr.backward(g)
    yo.backward(g.cuda())
    torch.cuda.synchronize()
    assert cosine(xo.grad, xr.grad) >= 0.95
    for (k, p), (_, q) in zip(ref.named_parameters(), ours.named_parameters()):
        if k.endswith("weight"):
            assert cosine(q.grad, p.grad) >= 0.95, k


def test_revgan_step_vs_oracle():
    from ganslate_b200.presets import revgan_vnet3d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle3d as O3
    from parity_util import cosine, rel_l2
    random.seed(0)
    oracle = O3.OracleRevGAN(O3.default_3d_conf(in_channels=2, out_channels=2, ndf=16, n_layers=2, **SMALL), seed=0)
    torch.manual_seed(0)
    ours = build_gan(revgan_vnet3d(channels=2, ndf=16, n_layers=2, **SMALL))
    for name in ("G", "D_B", "D_A"):
        _load(ours.networks[name], oracle.networks[name])
    a, b = O3.synthetic_volume(1, 2, 16, 32, seed=1)
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
            if k.endswith("weight") and p.grad is not None and p.grad.abs().max() > 0:
                c = cosine(pg[k].grad, p.grad)
                if c < 0.9:
                    bad.append((name, k, c))
    assert not bad, bad


def test_revgan_graph_replay_equals_eager_step():
    """RevGAN with train.cuda_graph (+ two-stream execution, inverse-recompute backward): a replayed iteration computes what
    an eager iteration computes from the same weights and inputs (losses within the run-to-run noise of the bf16 path)."""
    from ganslate_b200.presets import revgan_vnet3d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle3d as O3
    a, b = O3.synthetic_volume(1, 2, 16, 32, seed=1)
    torch.manual_seed(0)
    random.seed(0)
    mg = build_gan(revgan_vnet3d(channels=2, ndf=16, n_layers=2, cuda_graph=True, cuda_graph_warmup=2, multi_stream=True, **SMALL))
    for _ in range(4):   # 2 eager warm-up iterations, capture, one more replay
        mg.set_input({"A": a, "B": b})
        mg.optimize_parameters()
    torch.cuda.synchronize()
    assert len(mg._graphs) == 2 and mg.graph_launches_per_step > 100
    state = {n: {k: v.detach().clone() for k, v in net.state_dict().items()} for n, net in mg.networks.items()}
    mg.set_input({"A": a, "B": b})
    mg.optimize_parameters()
    torch.cuda.synchronize()
    lg = {k: float(v) for k, v in mg.losses.items() if v is not None}
    torch.manual_seed(0)
    me = build_gan(revgan_vnet3d(channels=2, ndf=16, n_layers=2, **SMALL))
    for n, net in me.networks.items():
        net.load_state_dict(state[n])
    me.set_input({"A": a, "B": b})
    me.optimize_parameters()
    torch.cuda.synchronize()
    le = {k: float(v) for k, v in me.losses.items() if v is not None}
    for k in le:
        assert abs(le[k] - lg[k]) <= 3e-2 * abs(le[k]) + 1e-3, (k, le[k], lg[k])
    w = next(iter(mg.networks["G"].parameters()))
    assert (w.detach() - next(iter(state["G"].values()))).abs().max().item() > 0   # the replay stepped the optimizer
