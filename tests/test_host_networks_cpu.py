"""Host logic of whole networks on the CPU: the modules of ganslate_b200.nn run forward + backward through
tests/fake_cabi.py (a pointer-level CPU restatement of the C ABI, test infrastructure only) and are compared with the
CPU oracle.  This checks everything ABOVE the ABI -- tape construction, fused-step selection, channel-slice views,
reflection / replicate borders, gradient routing and accumulation, weight packing specs -- for every network
family, including the ones whose CUDA path has not run on a B200 yet (Unet3D, Piresnet3D, separable V-Net).
It says nothing about the CUDA kernels: those are covered by the `-m gpu` parity tests.

Tolerances are the GPU tests' (bf16 storage points are reproduced by the fake backend): outputs relative L2 <= 3e-2,
gradients cosine >= 0.9 against the fp32 oracle."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fake_cabi  # noqa: E402
from parity_util import cosine, rel_l2  # noqa: E402


def _load(ours, oracle):
    assert list(ours.state_dict().keys()) == list(oracle.state_dict().keys())
    ours.load_state_dict(oracle.state_dict())


def _compare(ref, ours, x, call_ref=None, call_ours=None, out_tol=3e-2, cos_tol=0.9):
    call_ref = call_ref or (lambda m, t: m(t))
    call_ours = call_ours or (lambda m, t: m(t))
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yr, yo = call_ref(ref, xr), call_ours(ours, xo)
    assert yo.shape == yr.shape and rel_l2(yo, yr) <= out_tol, rel_l2(yo, yr)
    g = torch.randn_like(yr)
    ref.zero_grad()
    ours.zero_grad()
    yr.backward(g)
    yo.backward(g)
    assert cosine(xo.grad, xr.grad) >= cos_tol, cosine(xo.grad, xr.grad)
    pr, po = dict(ref.named_parameters()), dict(ours.named_parameters())
    bad = []
    for k, p in pr.items():
        if p.grad is None:
            assert po[k].grad is None or float(po[k].grad.abs().max()) == 0.0, k
            continue
        assert po[k].grad is not None, k
        if p.dim() > 1 and p.grad.abs().max() > 0:
            c = cosine(po[k].grad, p.grad)
            if c < cos_tol:
                bad.append((k, c))
    assert not bad, bad


def test_resnet2d_and_patchgan2d_host_logic(monkeypatch):
    lib = fake_cabi.install(monkeypatch)
    from ganslate_b200.nn.discriminators import PatchGAN2D
    from ganslate_b200.nn.generators import Resnet2D
    from oracle import torch_oracle as O
    torch.manual_seed(0)
    ref = O.init_weights(O.OracleResnet2D(3, 3, n_residual_blocks=2))
    ours = Resnet2D(3, 3, "instance", n_residual_blocks=2)
    _load(ours, ref)
    x = torch.rand(1, 3, 32, 32) * 2 - 1
    _compare(ref, ours, x)
    assert lib.calls["gb_conv_data"] > 0 and lib.calls["gb_conv_wgrad"] > 0 and lib.calls["gb_in_bwd"] > 0
    refd = O.init_weights(O.OraclePatchGAN2D(3, 16, 2))
    oursd = PatchGAN2D(3, 16, 2, (4, 4), "instance")
    _load(oursd, refd)
    _compare(refd, oursd, torch.rand(2, 3, 32, 32) * 2 - 1)


def test_unet2d_and_unet3d_host_logic(monkeypatch):
    fake_cabi.install(monkeypatch)
    from ganslate_b200.nn.generators import Unet2D, Unet3D
    from oracle import torch_oracle as O
    torch.manual_seed(1)
    ref = O.init_weights(O.OracleUnet2D(3, 2, 5, ngf=8))
    ours = Unet2D(3, 2, 5, "instance", ngf=8)
    _load(ours, ref)
    _compare(ref, ours, torch.rand(1, 3, 32, 64) * 2 - 1)
    ref3 = O.init_weights(O.OracleUnet3D(1, 1, 5, ngf=8))
    ours3 = Unet3D(1, 1, 5, "instance", ngf=8)
    _load(ours3, ref3)
    _compare(ref3, ours3, torch.rand(1, 1, 32, 32, 32) * 2 - 1)


SMALL = dict(first_layer_channels=8, down_blocks=(1, 1), up_blocks=(1, 1))


@pytest.mark.parametrize("separable", [False, True], ids=["dense", "separable"])
def test_vnet3d_host_logic(monkeypatch, separable):
    fake_cabi.install(monkeypatch)
    from ganslate_b200.nn.generators import Vnet3D
    from oracle import torch_oracle as O
    from oracle import torch_oracle3d as O3
    torch.manual_seed(0)
    ref = O.init_weights(O3.OracleVnet3D(1, 1, use_inverse=True, is_separable=separable, **SMALL))
    ours = Vnet3D(1, 1, "instance", use_memory_saving=False, use_inverse=True, is_separable=separable, **SMALL)
    _load(ours, ref)
    x, _ = O3.synthetic_volume(1, 1, 8, 16, seed=3)
    for inverse in (False, True):
        _compare(ref, ours, x, lambda m, t: m(t, inverse=inverse), lambda m, t: m(t, inverse=inverse))


def test_piresnet3d_and_patchgan3d_host_logic(monkeypatch):
    lib = fake_cabi.install(monkeypatch)
    from ganslate_b200.nn.discriminators import PatchGAN3D
    from ganslate_b200.nn.generators import Piresnet3D
    from oracle import torch_oracle as O
    from oracle import torch_oracle3d as O3
    torch.manual_seed(0)
    ref = O.init_weights(O3.OraclePiresnet3D(2, 2, 2, first_layer_channels=16, use_inverse=True))
    ours = Piresnet3D(2, 2, "instance", depth=2, first_layer_channels=16, use_memory_saving=False, use_inverse=True)
    _load(ours, ref)
    x, _ = O3.synthetic_volume(1, 2, 8, 12, seed=3)
    for inverse in (False, True):
        _compare(ref, ours, x, lambda m, t: m(t, inverse=inverse), lambda m, t: m(t, inverse=inverse))
    assert lib.calls["gb_replicate_pad_fwd"] > 0 and lib.calls["gb_replicate_pad_bwd"] > 0
    refd = O.init_weights(O3.OraclePatchGAN3D(1, 16, 2, (4, 4, 4)))
    oursd = PatchGAN3D(1, 16, 2, (4, 4, 4), "instance")
    _load(oursd, refd)
    xd, _ = O3.synthetic_volume(1, 1, 16, 16, seed=5)
    _compare(refd, oursd, xd)


def test_gradient_side_pixel_windows_host_logic(monkeypatch):
    """The opt-in gradient-side windows (ops.BWD_WINDOW_CONV) through the whole Resnet2D backward."""
    fake_cabi.install(monkeypatch)
    from ganslate_b200 import ops
    from ganslate_b200.nn.generators import Resnet2D
    from oracle import torch_oracle as O
    monkeypatch.setattr(ops, "BWD_WINDOW_CONV", True)
    torch.manual_seed(0)
    ref = O.init_weights(O.OracleResnet2D(3, 3, n_residual_blocks=1))
    ours = Resnet2D(3, 3, "instance", n_residual_blocks=1)
    _load(ours, ref)
    assert ours.model[-2].conv_op().bwd_window
    _compare(ref, ours, torch.rand(1, 3, 24, 24) * 2 - 1)


def test_resnet3d_slab_convolutions_host_logic(monkeypatch):
    """Resnet3D: the 7x7x7 layers (343 taps > GB_MAX_TAPS) run as seven (1,7,7) depth slabs accumulated in FP32
    (ops.SlabConv: depth-shifted views, per-slab packed weights read from weight[:, :, dz], per-slab weight-gradient
    unpack into dW[:, :, dz]); the first layer's slabs additionally use the pixel-window formulation."""
    lib = fake_cabi.install(monkeypatch)
    from ganslate_b200 import ops
    from ganslate_b200.nn.generators import Resnet3D
    from oracle import torch_oracle as O
    from oracle import torch_oracle3d as O3
    torch.manual_seed(0)
    ref = O.init_weights(O3.OracleResnet3D(1, 2, n_residual_blocks=1))
    ours = Resnet3D(1, 2, "instance", n_residual_blocks=1)
    _load(ours, ref)
    first, last = ours.model[1].conv_op(), ours.model[-2].conv_op()
    assert isinstance(first, ops.SlabConv) and isinstance(last, ops.SlabConv) and len(first.slabs) == 7
    assert first.slabs[0].window and not last.slabs[0].window
    x = torch.rand(1, 1, 8, 8, 8) * 2 - 1
    _compare(ref, ours, x)
    assert lib.calls["gb_conv_data"] >= 2 * 7


def test_slab_convolution_matches_torch_exactly_enough(monkeypatch):
    """One SlabConv layer against torch.nn.functional.conv3d on bf16-valued inputs: forward, data gradient, weight and
    bias gradient (tolerance 1e-2 max-relative, as tests/gpu_bringup.py uses on the GPU)."""
    fake_cabi.install(monkeypatch)
    from ganslate_b200.nn import layers
    from parity_util import max_rel
    torch.manual_seed(3)
    conv = layers.Conv3d(16, 24, (5, 6, 6), padding=(0, 1, 2), bias=True)   # 180 taps -> 5 slabs of 36
    with torch.no_grad():
        conv.weight.copy_((torch.randn_like(conv.weight) * 0.05).to(torch.bfloat16).float())
        conv.bias.copy_(torch.randn_like(conv.bias) * 0.1)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = torch.nn.Sequential(conv)

        def forward(self, t):
            return layers.run_network(self, list(self.model), t)

    x = torch.randn(2, 16, 7, 9, 8).to(torch.bfloat16).float().requires_grad_(True)
    y = Net()(x)
    xr = x.detach().clone().requires_grad_(True)
    wr, br = conv.weight.detach().clone().requires_grad_(True), conv.bias.detach().clone().requires_grad_(True)
    yr = torch.nn.functional.conv3d(xr, wr, br, padding=(0, 1, 2))
    g = torch.randn_like(yr).to(torch.bfloat16).float()
    y.backward(g)
    yr.backward(g)
    assert max_rel(y, yr) < 1e-2 and max_rel(x.grad, xr.grad) < 1e-2
    assert max_rel(conv.weight.grad, wr.grad) < 1e-2 and max_rel(conv.bias.grad, br.grad) < 1e-2


def test_cyclegan_iteration_host_logic(monkeypatch):
    """One whole CycleGAN iteration (ganslate_b200.nn.gans.unpaired.CycleGAN built by the plug-in builders: 4 generator
    passes, frozen-discriminator G step, two D steps with the ImagePool) through the fake backend against the CPU
    oracle's iteration -- the recipe's host logic without a GPU.  Optimizer steps are skipped on both sides (FusedAdam
    is a CUDA launch); losses, generated images and per-parameter gradients are compared."""
    import random
    fake_cabi.install(monkeypatch)
    from ganslate_b200.nn.gans import base
    from ganslate_b200.presets import cyclegan_resnet2d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle as O
    monkeypatch.setattr(base.BaseGAN, "_specify_device", lambda self: torch.device("cpu"))
    random.seed(0)
    oracle = O.OracleCycleGAN(O.default_cyclegan_conf(n_residual_blocks=1), seed=0)
    torch.manual_seed(0)
    ours = build_gan(cyclegan_resnet2d(batch_size=1, n_residual_blocks=1))
    for name in oracle.networks:  # same seed, same init order -> identical weights
        for (k1, p1), (k2, p2) in zip(oracle.networks[name].state_dict().items(), ours.networks[name].state_dict().items()):
            assert k1 == k2 and torch.equal(p1, p2), (name, k1)
    a, b = O.synthetic_batch(1, 3, 64, seed=1)  # (PatchGAN's last InstanceNorm sees 7x7 positions at this size)
    lo = oracle.optimize_parameters(a, b, step_optimizers=False)
    for o in ours.optimizers.values():
        monkeypatch.setattr(o, "step", lambda *args, **kw: None)
    ours.set_input({"A": a, "B": b})
    ours.optimize_parameters()
    for k, ref in lo.items():
        got = float(ours.losses[k].detach())
        assert abs(ref - got) <= 2e-2 * max(abs(ref), 1e-3), (k, ref, got)
    for k in ("fake_B", "rec_A", "fake_A", "rec_B"):
        assert rel_l2(ours.visuals[k], oracle.visuals[k]) < 8e-2, (k, rel_l2(ours.visuals[k], oracle.visuals[k]))
    bad = []
    for name in oracle.networks:
        po, pg = dict(oracle.networks[name].named_parameters()), dict(ours.networks[name].named_parameters())
        for k, p in po.items():
            if p.grad is None or p.dim() <= 1:
                continue
            c = cosine(pg[k].grad, p.grad)
            if c < 0.9:
                bad.append((name, k, c))
    assert not bad, bad


def test_revgan_piresnet3d_iteration_host_logic(monkeypatch):
    """The shipped BraTS experiment's pairing (RevGAN + Piresnet3D + PatchGAN3D(n_layers 2), revgan.yaml:25-39) at a
    small size: one whole iteration through the fake backend against the oracle's RevGAN iteration."""
    import random
    fake_cabi.install(monkeypatch)
    from ganslate_b200.nn.gans import base
    from ganslate_b200.presets import revgan_piresnet3d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle3d as O3
    monkeypatch.setattr(base.BaseGAN, "_specify_device", lambda self: torch.device("cpu"))
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
        monkeypatch.setattr(o, "step", lambda *args, **kw: None)
    ours.set_input({"A": a, "B": b})
    ours.optimize_parameters()
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


def test_revgan_graph_phases_follow_the_eager_order(monkeypatch):
    """RevGAN with `train.cuda_graph`: generator phase and discriminator phase are the captured units, the image pools are
    queried in between.  With the capture replaced by a direct call the phased path must run the eager program: same
    losses, same gradients, G stepped before the discriminators (revgan.py:89-116), pooled fakes handed to backward_D."""
    import random
    _cpu_recipe(monkeypatch)
    from ganslate_b200.presets import revgan_piresnet3d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle3d as O3
    a, b = O3.synthetic_volume(1, 1, 16, 32, seed=1)
    results = {}
    for mode in ("eager", "phases"):
        torch.manual_seed(0)
        random.seed(0)
        gan = build_gan(revgan_piresnet3d(channels=1, depth=2, first_layer_channels=16, ndf=16, n_layers=2))
        order = []
        for name, o in gan.optimizers.items():
            monkeypatch.setattr(o, "step", lambda *args, _n=name, **kw: order.append(_n))
        seen = []
        if mode == "phases":
            monkeypatch.setattr(gan, "graph_mode", lambda key: True)
            monkeypatch.setattr(gan, "run_graphed", lambda name, fn: (seen.append(name), fn())[1])
            monkeypatch.setattr(gan, "stage_input", lambda name, t: (seen.append(name), t)[1])
        gan.set_input({"A": a, "B": b})
        gan.optimize_parameters()
        results[mode] = ({k: float(v.detach()) for k, v in gan.losses.items() if v is not None},
                         {(n, k): p.grad.clone() for n, net in gan.networks.items() for k, p in net.named_parameters()
                          if p.grad is not None}, list(order), list(seen))
    assert results["phases"][3] == ["real_A", "real_B", "G", "pool_B", "pool_A", "D"]
    assert results["eager"][2] == results["phases"][2] == ["G", "D"]
    assert results["eager"][0] == results["phases"][0]
    assert results["eager"][1].keys() == results["phases"][1].keys()
    for k, v in results["eager"][1].items():
        assert torch.equal(v, results["phases"][1][k]), k


def test_bringup_cases_through_fake_backend(monkeypatch):
    """Every single-operator case of tests/gpu_bringup.py (the cases the GPU suite runs against torch: 1x1 ... 7x7,
    strided, transposed, 3-D convolutions incl. pixel windows; InstanceNorm groups with and without borders; residual
    blocks) through the fake backend on the CPU with the same tolerances -- this is what pins tests/fake_cabi.py to
    the semantics the GPU kernels are tested for."""
    fake_cabi.install(monkeypatch)
    import gpu_bringup
    monkeypatch.setattr(gpu_bringup, "dev", "cpu")
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    bad = [i for i, case in enumerate(gpu_bringup.CASES[:-1]) if not case()]  # (the last entry is the loss kernels)
    assert not bad, bad


def _cpu_recipe(monkeypatch):
    fake_cabi.install(monkeypatch)
    from ganslate_b200.nn.gans import base
    monkeypatch.setattr(base.BaseGAN, "_specify_device", lambda self: torch.device("cpu"))


def test_pix2pix_iteration_host_logic(monkeypatch):
    """Pix2PixConditionalGAN (paired recipe, discriminator on cat[A, B]) with the Resnet2D generator, one iteration
    through the fake backend against OraclePix2Pix (same cases as tests/test_pix2pix_gpu.py, smaller)."""
    _cpu_recipe(monkeypatch)
    from ganslate_b200.presets import pix2pix_resnet2d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle as O
    oracle = O.OraclePix2Pix(lambda_pix2pix=30.0, n_residual_blocks=1, n_layers=3, seed=0)
    torch.manual_seed(0)
    ours = build_gan(pix2pix_resnet2d(batch_size=1, n_layers=3, n_residual_blocks=1))
    for name in ("G", "D"):
        for (k1, p1), (k2, p2) in zip(oracle.networks[name].state_dict().items(), ours.networks[name].state_dict().items()):
            assert k1 == k2 and torch.equal(p1, p2), (name, k1)
    a, b = O.synthetic_batch(1, 3, 64, seed=1, width=96)
    lo = oracle.optimize_parameters(a, b, step_optimizers=False)
    for o in ours.optimizers.values():
        monkeypatch.setattr(o, "step", lambda *args, **kw: None)
    ours.set_input({"A": a, "B": b})
    ours.optimize_parameters()
    for k, v in lo.items():
        assert abs(float(ours.losses[k].detach()) - v) <= 2e-2 * abs(v), (k, v, float(ours.losses[k].detach()))
    assert rel_l2(ours.visuals["fake_B"], oracle.visuals["fake_B"]) < 3e-2
    for name in ("G", "D"):
        po, pg = dict(oracle.networks[name].named_parameters()), dict(ours.networks[name].named_parameters())
        for k in po:
            if k.endswith("weight") and po[k].dim() > 1:
                assert cosine(pg[k].grad, po[k].grad) > 0.9, (name, k, cosine(pg[k].grad, po[k].grad))


def test_cut_iteration_host_logic(monkeypatch):
    """CUT (PatchNCE over encoder feature taps + FeaturePatchMLP), one iteration through the fake backend against
    OracleCUT with injected patch ids (same protocol as tests/test_cut_gpu.py, smaller)."""
    _cpu_recipe(monkeypatch)
    from ganslate_b200.presets import cut_resnet2d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle as O
    torch.manual_seed(0)
    oracle = O.OracleCUT(n_residual_blocks=9, num_patches=64, seed=0)
    torch.manual_seed(0)
    conf = cut_resnet2d()
    conf.train.gan.optimizer.num_patches = 64
    ours = build_gan(conf)
    for name in ("G", "D", "mlp"):
        for (k1, p1), (k2, p2) in zip(oracle.networks[name].state_dict().items(), ours.networks[name].state_dict().items()):
            assert k1 == k2 and torch.equal(p1, p2), (name, k1)
    a, b = O.synthetic_batch(1, 3, 64, seed=1)
    g = torch.Generator().manual_seed(5)
    sizes = [70 * 70, 32 * 32, 16 * 16, 16 * 16, 16 * 16]
    ids = [torch.randperm(s, generator=g)[:64] for s in sizes]
    lo, _ = oracle.optimize_parameters(a, b, patch_ids=ids, step_optimizers=False)
    ours.fixed_patch_ids = ids
    for o in ours.optimizers.values():
        monkeypatch.setattr(o, "step", lambda *args, **kw: None)
    ours.set_input({"A": a, "B": b})
    ours.optimize_parameters()
    for k, v in lo.items():
        assert abs(float(ours.losses[k].detach()) - v) <= 2e-2 * abs(v), (k, v, float(ours.losses[k].detach()))
    for name in ("mlp", "G", "D"):
        po, pg = dict(oracle.networks[name].named_parameters()), dict(ours.networks[name].named_parameters())
        for k in po:
            if k.endswith("weight") and po[k].grad is not None and po[k].dim() > 1:
                assert cosine(pg[k].grad, po[k].grad) > 0.85, (name, k, cosine(pg[k].grad, po[k].grad))


def _random_conv_cases(n, seed):
    import random
    rng = random.Random(seed)
    cases = []
    while len(cases) < n:
        dims = rng.choice([2, 2, 3])
        transposed = rng.random() < 0.35
        k = tuple(rng.choice([1, 2, 3, 4, 5]) for _ in range(dims))
        s = tuple(rng.choice([1, 1, 2, 3]) for _ in range(dims))
        p = tuple(rng.randint(0, max(0, kk - 1) // 1 if kk > 1 else 0) for kk in k)
        p = tuple(min(pp, kk - 1) for pp, kk in zip(p, k))
        op = tuple(rng.randint(0, ss - 1) for ss in s) if transposed else (0,) * dims
        cin, cout = rng.choice([1, 3, 8, 12, 16, 40, 64]), rng.choice([1, 3, 8, 24, 64, 72])
        ext = tuple(rng.randint(max(kk, 2), 9) for kk in k)
        ntaps = 1
        for kk in k:
            ntaps *= kk
        if ntaps > 64 or (transposed and any(pp > kk - 1 for pp, kk in zip(p, k))):
            continue
        cases.append(dict(dims=dims, transposed=transposed, k=k, s=s, p=p, op=op, cin=cin, cout=cout, ext=ext,
                          N=rng.choice([1, 2]), bias=rng.random() < 0.7))
    return cases


@pytest.mark.parametrize("case", _random_conv_cases(36, 2024), ids=lambda c: "{}{}d k{} s{} p{} {}->{}".format(
    "T" if c["transposed"] else "", c["dims"], "x".join(map(str, c["k"])), "x".join(map(str, c["s"])),
    "x".join(map(str, c["p"])), c["cin"], c["cout"]))
def test_random_convolutions_through_the_host_path(monkeypatch, case):
    """Randomly drawn (transposed) convolutions, 2-D and 3-D, odd channel counts, strides up to 3, output padding:
    forward, data gradient, weight and bias gradient of layers.Conv* through ConvOp's class / tap specs, packing and
    weight-gradient plans (fake backend) against torch -- the geometry cases no network of the suite happens to use."""
    fake_cabi.install(monkeypatch)
    from ganslate_b200.nn import layers
    from parity_util import max_rel
    torch.manual_seed(7)
    c = case
    if c["transposed"]:
        cls = layers.ConvTranspose3d if c["dims"] == 3 else layers.ConvTranspose2d
        conv = cls(c["cin"], c["cout"], c["k"], stride=c["s"], padding=c["p"], output_padding=c["op"], bias=c["bias"])
        fn = torch.nn.functional.conv_transpose3d if c["dims"] == 3 else torch.nn.functional.conv_transpose2d
        kw = dict(stride=c["s"], padding=c["p"], output_padding=c["op"])
    else:
        cls = layers.Conv3d if c["dims"] == 3 else layers.Conv2d
        conv = cls(c["cin"], c["cout"], c["k"], stride=c["s"], padding=c["p"], bias=c["bias"])
        fn = torch.nn.functional.conv3d if c["dims"] == 3 else torch.nn.functional.conv2d
        kw = dict(stride=c["s"], padding=c["p"])
    with torch.no_grad():
        conv.weight.copy_((torch.randn_like(conv.weight) * 0.1).to(torch.bfloat16).float())
        if c["bias"]:
            conv.bias.copy_(torch.randn_like(conv.bias) * 0.1)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = torch.nn.Sequential(conv)

        def forward(self, t):
            return layers.run_network(self, list(self.model), t)

    nclass = 1
    for ss in c["s"]:
        nclass *= ss
    if nclass > 8:  # documented limit (GB_MAX_CLASSES parity classes): refused loudly, never computed wrongly
        with pytest.raises(ValueError, match="parity classes"):
            conv.conv_op()
        return
    x = torch.randn((c["N"], c["cin"]) + c["ext"]).to(torch.bfloat16).float().requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    wr = conv.weight.detach().clone().requires_grad_(True)
    br = conv.bias.detach().clone().requires_grad_(True) if c["bias"] else None
    yr = fn(xr, wr, br, **kw)
    if min(yr.shape) == 0:
        pytest.skip("empty output")
    y = Net()(x)
    assert y.shape == yr.shape
    g = torch.randn_like(yr).to(torch.bfloat16).float()
    y.backward(g)
    yr.backward(g)
    assert max_rel(y, yr) < 1e-2, max_rel(y, yr)
    assert max_rel(x.grad, xr.grad) < 1e-2, max_rel(x.grad, xr.grad)
    assert max_rel(conv.weight.grad, wr.grad) < 1e-2, max_rel(conv.weight.grad, wr.grad)
    if c["bias"]:
        assert max_rel(conv.bias.grad, br.grad) < 1e-2


def test_direct_parameter_gradients_equal_autograd_accumulation(monkeypatch):
    """ops.DIRECT_PARAM_GRAD (opt-in): the tape writes / accumulates parameter gradients in `param.grad` itself
    (unpack accumulate flag, bias-gradient sums added in place) instead of returning them to autograd.  One whole
    CycleGAN iteration -- every generator is used twice in backward_G, every discriminator sees a real and a fake batch
    -- must leave the same gradients in both modes, and no gradient may reach autograd's AccumulateGrad in direct mode."""
    import random
    fake_cabi.install(monkeypatch)
    from ganslate_b200 import ops
    from ganslate_b200.nn.gans import base
    from ganslate_b200.presets import cyclegan_resnet2d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle as O
    monkeypatch.setattr(base.BaseGAN, "_specify_device", lambda self: torch.device("cpu"))
    a, b = O.synthetic_batch(1, 3, 64, seed=1)
    grads, hooks = {}, {}
    for direct in (False, True):
        monkeypatch.setattr(ops, "DIRECT_PARAM_GRAD", direct)
        random.seed(0)
        torch.manual_seed(0)
        gan = build_gan(cyclegan_resnet2d(batch_size=1, n_residual_blocks=1))
        for o in gan.optimizers.values():
            monkeypatch.setattr(o, "step", lambda *args, **kw: None)
        fired = []
        for name, net in gan.networks.items():
            for k, p in net.named_parameters():
                p.register_hook(lambda g, key=(name, k): fired.append(key) if g is not None else None)
        gan.set_input({"A": a, "B": b})
        gan.optimize_parameters()
        grads[direct] = {(n, k): p.grad.clone() for n, net in gan.networks.items() for k, p in net.named_parameters()}
        hooks[direct] = len(fired)
    assert hooks[False] > 0 and hooks[True] == 0
    assert grads[False].keys() == grads[True].keys()
    for key, g in grads[False].items():
        d = grads[True][key]
        assert d.shape == g.shape
        assert torch.allclose(d, g, rtol=1e-5, atol=1e-6 * max(1.0, float(g.abs().max()))), (key, float((d - g).abs().max()))


def test_cut_graph_segments_follow_the_eager_order(monkeypatch):
    """CUT with `train.cuda_graph`: the iteration is split into the segments that are captured (forward + D phase, D
    step, G + patch-MLP phase, their steps).  With the capture itself replaced by a direct call, the segmented path must
    run the same program as the eager one: same losses, same gradients, optimizers stepped in the reference's order
    (D before the generator phase: cut.py:121-125)."""
    _cpu_recipe(monkeypatch)
    from ganslate_b200.nn.gans import base
    from ganslate_b200.presets import cut_resnet2d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle as O
    a, b = O.synthetic_batch(1, 3, 64, seed=1)
    g = torch.Generator().manual_seed(5)
    ids = [torch.randperm(s, generator=g)[:64] for s in [70 * 70, 32 * 32, 16 * 16, 16 * 16, 16 * 16]]
    results = {}
    for mode in ("eager", "segments"):
        torch.manual_seed(0)
        conf = cut_resnet2d()
        conf.train.gan.optimizer.num_patches = 64
        gan = build_gan(conf)
        order = []
        for name, o in gan.optimizers.items():
            monkeypatch.setattr(o, "step", lambda *args, _n=name, **kw: order.append(_n))
        if mode == "segments":
            seen = []
            monkeypatch.setattr(gan, "graph_mode", lambda key: True)
            monkeypatch.setattr(gan, "run_graphed", lambda name, fn: (seen.append(name), fn())[1])
        gan.fixed_patch_ids = ids
        gan.set_input({"A": a, "B": b})
        gan.optimize_parameters()
        results[mode] = ({k: float(v.detach()) for k, v in gan.losses.items() if v is not None},
                         {(n, k): p.grad.clone() for n, net in gan.networks.items() for k, p in net.named_parameters()
                          if p.grad is not None}, list(order))
    assert seen == ["D", "G"]
    assert results["eager"][2] == results["segments"][2] == ["D", "G", "mlp"]
    assert results["eager"][0] == results["segments"][0]
    assert results["eager"][1].keys() == results["segments"][1].keys()
    for k, v in results["eager"][1].items():
        assert torch.equal(v, results["segments"][1][k]), k


@pytest.mark.parametrize("inverse", [False, True], ids=["forward", "inverse"])
def test_vnet3d_memory_saving_recompute_matches_kept_activations(monkeypatch, inverse):
    """use_memory_saving=True (ganslate/nn/invertible.py:8-48, vnet3d.py:36-52): coupling inputs are freed in the forward
    pass and rebuilt by the inverse coupling in backward.  Same output bit for bit; gradients equal those of the
    keep-everything mode up to the bf16 rounding of the rebuilt inputs (stated bound: relative L2 <= 5e-2 per tensor --
    measured 0 - 3.9e-2, against the 0.1 - 0.5 both modes are away from the fp32 oracle at this depth -- and no further
    from the fp32 oracle than the kept mode + 2e-2)."""
    fake_cabi.install(monkeypatch)
    from ganslate_b200.nn import invertible
    from ganslate_b200.nn.generators import Vnet3D
    from oracle import torch_oracle as O
    from oracle import torch_oracle3d as O3
    torch.manual_seed(0)
    cfg = dict(first_layer_channels=8, down_blocks=(1, 2, 3), up_blocks=(3, 2, 1))
    ref = O.init_weights(O3.OracleVnet3D(1, 1, use_inverse=True, **cfg))
    keep = Vnet3D(1, 1, "instance", use_memory_saving=False, use_inverse=True, **cfg)
    save = Vnet3D(1, 1, "instance", use_memory_saving=True, use_inverse=True, **cfg)
    _load(keep, ref)
    _load(save, ref)
    x, _ = O3.synthetic_volume(1, 1, 8, 16, seed=3)
    outs, grads = {}, {}
    g = None
    for name, net in (("ref", ref), ("keep", keep), ("save", save)):
        xi = x.clone().requires_grad_(True)
        before = dict(invertible.RECOMPUTE_STATS)
        y = net(xi, inverse=inverse)
        g = torch.randn_like(y) if g is None else g
        net.zero_grad()
        y.backward(g)
        outs[name] = y.detach()
        grads[name] = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        grads[name]["__input__"] = xi.grad.clone()
        ran = {k: invertible.RECOMPUTE_STATS[k] - before[k] for k in before}
        if name == "save":
            # every coupling block re-runs in backward; all but the first block of each sequence rebuild their input
            assert ran == {"blocks": 12, "rebuilt_inputs": 6}, ran
        else:
            assert ran == {"blocks": 0, "rebuilt_inputs": 0}, ran
    assert torch.equal(outs["keep"], outs["save"])
    worst = 0.0
    for k, gk in grads["keep"].items():
        if gk.dim() > 1 and gk.abs().max() > 0:
            d = rel_l2(grads["save"][k], gk)
            worst = max(worst, d)
            assert d <= 5e-2, (k, d)
            assert rel_l2(grads["save"][k], grads["ref"][k]) <= rel_l2(gk, grads["ref"][k]) + 2e-2, k
    assert 0.0 < worst  # the rebuilt inputs really differ by a rounding: the path ran on different bits
    assert set(grads["save"]) == set(grads["keep"])


def test_memory_saving_frees_coupling_inputs_between_forward_and_backward(monkeypatch):
    """Between forward and backward only the first block's input and the last block's output of a coupling sequence
    hold memory; the other block inputs are released (and every intermediate of the blocks was never recorded)."""
    fake_cabi.install(monkeypatch)
    from ganslate_b200.nn import invertible, layers
    from ganslate_b200.nn.generators import Vnet3D
    torch.manual_seed(0)
    net = Vnet3D(1, 1, "instance", use_memory_saving=True, use_inverse=False, first_layer_channels=8,
                 down_blocks=(3,), up_blocks=(3,))
    seen = []
    orig = invertible.recompute_coupling

    def spy(tape, x, fn, inverse, free_input):
        y = orig(tape, x, fn, inverse, free_input)
        seen.append((x.st, free_input))
        return y

    monkeypatch.setattr(invertible, "recompute_coupling", spy)
    x = torch.rand(1, 1, 8, 16, 16) * 2 - 1
    y = net(x)
    assert [f for _, f in seen] == [False, True, True] * 2
    assert all((st.t is None) == f for st, f in seen)
    n_steps = len(y.grad_fn.tape.steps)
    y.sum().backward()
    assert all(st.t is not None for st, _ in seen)   # rebuilt during backward
    # and the recorded tape is what a network without the couplings' inner steps would record
    keep = Vnet3D(1, 1, "instance", use_memory_saving=False, use_inverse=True, first_layer_channels=8,
                  down_blocks=(3,), up_blocks=(3,))
    assert len(keep(x).grad_fn.tape.steps) > n_steps


def test_tape_is_released_after_backward_and_retain_graph_mode(monkeypatch):
    """no activation / gradient buffer outlives its backward; a second backward raises like torch's
    does; with layers.RELEASE_TAPE off a retain_graph=True second backward gives the same gradients again (no stale
    accumulation)."""
    fake_cabi.install(monkeypatch)
    from ganslate_b200.nn import layers
    from ganslate_b200.nn.generators import Resnet2D
    torch.manual_seed(0)
    net = Resnet2D(3, 3, "instance", n_residual_blocks=1)
    from ganslate_b200.nn.utils import init_weights
    init_weights(net, "normal", 0.02)
    x = (torch.rand(1, 3, 16, 16) * 2 - 1).requires_grad_(True)
    y = net(x)
    fn = y.grad_fn
    y.sum().backward(retain_graph=True)
    assert fn.tape is None and fn.b0 is None and not layers._LIVE_GRADS
    with pytest.raises(RuntimeError, match="second time"):
        y.sum().backward()
    monkeypatch.setattr(layers, "RELEASE_TAPE", False)
    net.zero_grad()
    x.grad = None
    y = net(x)
    y.sum().backward(retain_graph=True)
    g1 = {k: p.grad.clone() for k, p in net.named_parameters()}
    gx1 = x.grad.clone()
    net.zero_grad()
    x.grad = None
    y.sum().backward()
    assert torch.allclose(x.grad, gx1, rtol=1e-5, atol=1e-7)
    for k, p in net.named_parameters():
        assert torch.allclose(p.grad, g1[k], rtol=1e-4, atol=1e-7), k


def test_teacher_forced_parity_harness_on_the_fake_backend(monkeypatch):
    """tests/forced_parity.py (the harness of tests/test_forced_parity_gpu.py) through the fake backend: with forcing
    every layer agrees with the bf16-point oracle to rounding level and the weight gradients to summation order;
    without forcing the same run drifts by orders of magnitude more (the harness measures what it claims)."""
    fake_cabi.install(monkeypatch)
    from forced_parity import forced_network_parity, summarize
    from ganslate_b200.nn.discriminators import PatchGAN2D, PatchGAN3D
    from ganslate_b200.nn.generators import Resnet2D
    from oracle import torch_oracle as O
    from oracle import torch_oracle3d as O3
    torch.manual_seed(0)
    ref = O.init_weights(O.OracleResnet2D(3, 3, 2))
    ours = Resnet2D(3, 3, "instance", 2)
    _load(ours, ref)
    x = torch.rand(1, 3, 32, 32) * 2 - 1
    forced = summarize(forced_network_parity(ours, ref, x))
    free = summarize(forced_network_parity(ours, ref, x, force=False))
    assert forced["fwd_max_rel"] <= 1e-2 and forced["bwd_max_rel"] <= 1e-2 and forced["wgrad_max_rel"] <= 1e-4, forced
    assert free["bwd_rel_l2"] > 20 * forced["bwd_rel_l2"] and free["wgrad_rel_l2"] > 1e-2, (free, forced)
    assert forced["bwd_outlier_frac"] <= 1e-5 and free["bwd_outlier_frac"] > 1e-3, (free, forced)
    refd = O.init_weights(O.OraclePatchGAN2D(3, 16, 2))
    oursd = PatchGAN2D(3, 16, 2, (4, 4), "instance")
    _load(oursd, refd)
    s = summarize(forced_network_parity(oursd, refd, torch.rand(2, 3, 32, 32) * 2 - 1))
    assert s["fwd_max_rel"] <= 1e-2 and s["bwd_max_rel"] <= 1e-2 and s["wgrad_max_rel"] <= 1e-4, s
    ref3 = O.init_weights(O3.OraclePatchGAN3D(1, 16, 2, (4, 4, 4)))
    ours3 = PatchGAN3D(1, 16, 2, (4, 4, 4), "instance")
    _load(ours3, ref3)
    x3, _ = O3.synthetic_volume(1, 1, 16, 16, seed=5)
    s = summarize(forced_network_parity(ours3, ref3, x3))
    assert s["fwd_max_rel"] <= 1e-2 and s["bwd_max_rel"] <= 1e-2 and s["wgrad_max_rel"] <= 1e-4, s


def _fp32_compare(ref, ours, x, call=None, tol=2e-5):
    call = call or (lambda m, t: m(t))
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yr, yo = call(ref, xr), call(ours, xo)
    from parity_util import max_rel
    assert yo.shape == yr.shape and max_rel(yo, yr) <= tol, max_rel(yo, yr)
    g = torch.randn(yr.shape, generator=torch.Generator().manual_seed(9))
    ref.zero_grad()
    ours.zero_grad()
    yr.backward(g)
    yo.backward(g)
    assert max_rel(xo.grad, xr.grad) <= tol, max_rel(xo.grad, xr.grad)
    po = dict(ours.named_parameters())
    wmax = max(p.grad.abs().max().item() for k, p in ref.named_parameters() if p.grad is not None and p.dim() > 1)
    for k, p in ref.named_parameters():
        if p.grad is None:
            continue
        if p.dim() > 1 or p.grad.abs().max() > 1e-3 * wmax:
            assert max_rel(po[k].grad, p.grad) <= tol, (k, max_rel(po[k].grad, p.grad))
        else:   # bias in front of an InstanceNorm: mathematically zero
            assert (po[k].grad - p.grad).abs().max().item() <= 1e-4 * wmax, k


def test_fp32_validation_mode_matches_the_fp32_oracle_to_1e_5(monkeypatch):
    """ops.FP32_MODE (nn/fp32_mode.py): 3-way bf16-split operands through the same conv entry points (six launches per
    product, fp32 accumulation), fp32 buffers, fp32 norm / activation steps.  Through the fake backend the whole
    forward + backward of every network family agrees with the fp32 oracle to 2e-5 max-relative -- three orders of
    magnitude below the bf16 path, so one switch separates a kernel / wiring bug from rounding noise."""
    lib = fake_cabi.install(monkeypatch)
    from ganslate_b200 import ops
    from ganslate_b200.nn.discriminators import PatchGAN2D, PatchGAN3D
    from ganslate_b200.nn.generators import Resnet2D, Unet2D, Vnet3D
    from oracle import torch_oracle as O
    from oracle import torch_oracle3d as O3
    monkeypatch.setattr(ops, "FP32_MODE", True)
    torch.manual_seed(0)
    ref = O.init_weights(O.OracleResnet2D(3, 3, n_residual_blocks=2))
    ours = Resnet2D(3, 3, "instance", n_residual_blocks=2)
    _load(ours, ref)
    n0 = lib.calls.get("gb_conv_data", 0)
    _fp32_compare(ref, ours, torch.rand(1, 3, 32, 32) * 2 - 1)
    assert lib.calls["gb_conv_data"] - n0 >= 6 * 2 * 9      # six split products per convolution and direction
    refd = O.init_weights(O.OraclePatchGAN2D(3, 16, 2))
    oursd = PatchGAN2D(3, 16, 2, (4, 4), "instance")
    _load(oursd, refd)
    _fp32_compare(refd, oursd, torch.rand(2, 3, 32, 32) * 2 - 1)
    refu = O.init_weights(O.OracleUnet2D(3, 2, 5, ngf=8))
    oursu = Unet2D(3, 2, 5, "instance", ngf=8)
    _load(oursu, refu)
    _fp32_compare(refu, oursu, torch.rand(1, 3, 32, 64) * 2 - 1)
    refv = O.init_weights(O3.OracleVnet3D(1, 1, use_inverse=True, **SMALL))
    oursv = Vnet3D(1, 1, "instance", use_memory_saving=False, use_inverse=True, **SMALL)
    _load(oursv, refv)
    xv, _ = O3.synthetic_volume(1, 1, 8, 16, seed=3)
    for inverse in (False, True):
        _fp32_compare(refv, oursv, xv, call=lambda m, t: m(t, inverse=inverse), tol=5e-5)
    ref3 = O.init_weights(O3.OraclePatchGAN3D(1, 16, 2, (4, 4, 4)))
    ours3 = PatchGAN3D(1, 16, 2, (4, 4, 4), "instance")
    _load(ours3, ref3)
    x3, _ = O3.synthetic_volume(1, 1, 16, 16, seed=5)
    _fp32_compare(ref3, ours3, x3)
    # and teacher-forced against the oracle walked without rounding (tests/test_fp32_mode_gpu.py does this on the GPU)
    from forced_parity import forced_network_parity, summarize
    sm = summarize(forced_network_parity(ours, ref, torch.rand(1, 3, 32, 32) * 2 - 1, fp32=True))
    assert sm["fwd_max_rel"] <= 2e-5 and sm["bwd_max_rel"] <= 2e-5 and sm["wgrad_max_rel"] <= 2e-5, sm


def test_widened_few_channel_input_through_fake_backend(monkeypatch):
    """ops.WIDEN_INPUT: a 1 -> 16 channel 5x5x5 layer reads a 16-channel copy of its 8-channel operand (packing, weight-
    gradient plan and unpack then work on cin_pad = 16, the data gradient still writes the 8-channel buffer).  Same
    outputs and gradients as torch, and as the same layer with the switch off."""
    fake_cabi.install(monkeypatch)
    import gpu_bringup
    from ganslate_b200 import ops
    monkeypatch.setattr(gpu_bringup, "dev", "cpu")
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    for widen in (True, False):
        monkeypatch.setattr(ops, "WIDEN_INPUT", widen)
        op = ops.ConvOp(1, 16, (5, 5, 5), (1, 1, 1), (2, 2, 2))
        assert op.widen_input == widen and op.cin_pad == (16 if widen else 8)
        assert gpu_bringup.conv_case(f"3d k5 p2 1->16 4x12x10 widen={widen}", 1, 16, 5, 1, 2, 12, 10, D=4)
        assert gpu_bringup.conv_case(f"3d k5 p2 4->16 3x9x11 N=2 widen={widen}", 4, 16, 5, 1, 2, 9, 11, N=2, D=3)
    # not widened: 2-D kernels, strided layers, wide outputs, depth slabs of a larger kernel
    monkeypatch.setattr(ops, "WIDEN_INPUT", True)
    assert not ops.ConvOp(3, 16, (1, 7, 7), (1, 1, 1), (0, 3, 3)).widen_input
    assert not ops.ConvOp(1, 16, (4, 4, 4), (2, 2, 2), (1, 1, 1)).widen_input
    assert not ops.ConvOp(1, 64, (5, 5, 5), (1, 1, 1), (2, 2, 2)).widen_input
    assert not ops.ConvOp(1, 16, (1, 7, 7), (1, 1, 1), (0, 3, 3), weight_taps=343).widen_input


def test_other_batch_shape_after_capture_host_logic(monkeypatch):
    """BaseGAN.stage_input / graph_mode: a batch whose shape differs from the captured one runs ONE eager iteration on the
    tensor itself; the entries of visuals / losses the graphs were bound to are restored before the next replay."""
    _cpu_recipe(monkeypatch)
    from ganslate_b200.presets import cyclegan_resnet2d
    from ganslate_b200.utils.builders import build_gan
    gan = build_gan(cyclegan_resnet2d(n_residual_blocks=1))
    gan.use_cuda_graph, gan.graph_warmup_iters, gan._graph_calls = True, 0, 5
    a = torch.zeros(2, 3, 32, 32)
    buf = gan.stage_input("real_A", a)               # before any capture: allocates the static buffer
    assert buf is gan._static["real_A"] and buf.shape == a.shape
    gan._graphs = {"G": object(), "D": object()}     # "captured"
    bound_fake, bound_loss = torch.ones(1), torch.ones(())
    gan.visuals["fake_B"], gan.losses["G_AB"] = bound_fake, bound_loss
    small = torch.zeros(1, 3, 32, 32)
    got = gan.stage_input("real_A", small)           # partial batch
    assert got.shape == small.shape and gan._static["real_A"].shape == a.shape
    assert gan.graph_mode("step") is False           # this iteration is eager ...
    gan.visuals["fake_B"], gan.losses["G_AB"] = torch.zeros(1), torch.zeros(())   # ... and rebinds the entries
    assert gan.stage_input("real_A", a) is gan._static["real_A"]
    assert gan.graph_mode("step") is True            # next full batch: graphs again
    assert gan.visuals["fake_B"] is bound_fake and gan.losses["G_AB"] is bound_loss
