"""Pins the CPU oracle: against the golden fixtures generated from the reference's own modules
(oracle/make_golden.py) and, when /root/reference is present, against those modules directly."""
import json
import os
import random

import pytest
import torch

from oracle import reference_import as R
from oracle import torch_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def digest(t):
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, 5).long()
    return {"sum": t.sum().item(), "abs_sum": t.abs().sum().item(), "sq_sum": (t * t).sum().item(),
            "samples": t[idx].tolist(), "numel": t.numel()}


def close(a, b, rtol=2e-4, atol=1e-7):
    return abs(a - b) <= atol + rtol * max(abs(a), abs(b))


def assert_digest(got, ref, what, rtol=2e-4, atol=1e-6):
    assert got["numel"] == ref["numel"], what
    # sums of signed values cancel: compare against the magnitude scale (abs_sum)
    assert abs(got["sum"] - ref["sum"]) <= atol + rtol * ref["abs_sum"], (what, got["sum"], ref["sum"])
    assert close(got["abs_sum"], ref["abs_sum"], rtol, atol), (what, got["abs_sum"], ref["abs_sum"])
    assert close(got["sq_sum"], ref["sq_sum"], 2 * rtol, atol), (what, got["sq_sum"], ref["sq_sum"])
    scale = (ref["sq_sum"] / ref["numel"])**0.5
    for g, r in zip(got["samples"], ref["samples"]):
        assert abs(g - r) <= atol + 5 * rtol * max(scale, abs(r)), (what, g, r)


@pytest.mark.parametrize("name", ["cyclegan_step_32px_2blk", "cyclegan_step_64px_3blk_idt"])
def test_oracle_matches_reference_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        gold = json.load(f)
    c = gold["config"]
    random.seed(0)
    m = O.OracleCycleGAN(O.default_cyclegan_conf(n_residual_blocks=c["n_blocks"], lambda_identity=c["lambda_identity"]),
                         seed=c["seed"])
    for n, keys in gold["state_dict_keys"].items():
        assert list(m.networks[n].state_dict().keys()) == keys
        assert sum(p.numel() for p in m.networks[n].parameters()) == gold["param_counts"][n]
    a, b = O.synthetic_batch(c["batch"], 3, c["size"], seed=c["data_seed"])
    # snapshot the G gradients before the D phase exactly like the generator of the fixture did
    losses = {}
    m.visuals["real_A"], m.visuals["real_B"] = a, b
    ds = [m.networks["D_B"], m.networks["D_A"]]
    m.forward()
    m._set_requires_grad(ds, False)
    m.optimizers["G"].zero_grad(set_to_none=True)
    m.backward_G()
    grads = {f"{n}.{k}": digest(p.grad) for n in ("G_AB", "G_BA") for k, p in m.networks[n].named_parameters()}
    m.optimizers["G"].step()
    m._set_requires_grad(ds, True)
    m.optimizers["D"].zero_grad(set_to_none=True)
    m.backward_D("D_B")
    m.backward_D("D_A")
    grads.update({f"{n}.{k}": digest(p.grad) for n in ("D_B", "D_A") for k, p in m.networks[n].named_parameters()})
    m.optimizers["D"].step()
    losses = {k: float(v) for k, v in m.losses.items() if v is not None}
    for k, v in gold["losses"].items():
        assert close(losses[k], v, 1e-5), (k, losses[k], v)
    for k, d in gold["visuals"].items():
        assert_digest(digest(m.visuals[k]), d, k)
    for k, d in gold["grads"].items():
        assert_digest(grads[k], d, k, rtol=1e-3)
    for n in m.networks:
        for k, p in m.networks[n].named_parameters():
            assert_digest(digest(p), gold["weights_after_step"][f"{n}.{k}"], f"weight {n}.{k}")


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
def test_oracle_networks_equal_reference_modules():
    m = R.modules()
    torch.manual_seed(3)
    ref_g = m["Resnet2D"](3, 3, "instance", 2)
    m["init_weights"](ref_g, "normal", 0.02)
    torch.manual_seed(3)
    ora_g = O.init_weights(O.OracleResnet2D(3, 3, 2))
    assert list(ref_g.state_dict().keys()) == list(ora_g.state_dict().keys())
    for (k, a), (_, b) in zip(ref_g.state_dict().items(), ora_g.state_dict().items()):
        assert torch.equal(a, b), k
    torch.manual_seed(4)
    ref_d = m["PatchGAN2D"](3, 64, 3, (4, 4), "instance")
    m["init_weights"](ref_d, "normal", 0.02)
    torch.manual_seed(4)
    ora_d = O.init_weights(O.OraclePatchGAN2D(3, 64, 3, (4, 4)))
    for (k, a), (_, b) in zip(ref_d.state_dict().items(), ora_d.state_dict().items()):
        assert torch.equal(a, b), k
    x = torch.rand(2, 3, 48, 48) * 2 - 1
    yg, yo = ref_g(x), ora_g(x)
    assert torch.allclose(yg, yo, atol=1e-6)
    assert torch.allclose(ref_d(yg), ora_d(yo), atol=1e-6)
    crit = m["AdversarialLoss"]("lsgan")
    for flag in (True, False):
        assert torch.allclose(crit(ref_d(yg), target_is_real=flag), O.adversarial_lsgan(ora_d(yo), flag), atol=1e-7)


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
def test_b200_modules_mirror_reference_state_dict():
    """Drop-in contract: same state_dict keys, parameter order, shapes and same-seed initial weights."""
    from ganslate_b200.nn.discriminators import PatchGAN2D
    from ganslate_b200.nn.generators import Resnet2D
    from ganslate_b200.nn.utils import init_weights
    m = R.modules()
    for build_ref, build_ours in (
        (lambda: m["Resnet2D"](3, 3, "instance", 9), lambda: Resnet2D(3, 3, "instance", 9)),
        (lambda: m["PatchGAN2D"](3, 64, 3, (4, 4), "instance"), lambda: PatchGAN2D(3, 64, 3, (4, 4), "instance")),
    ):
        torch.manual_seed(11)
        ref = build_ref()
        m["init_weights"](ref, "normal", 0.02)
        torch.manual_seed(11)
        ours = build_ours()
        init_weights(ours, "normal", 0.02)
        assert list(ref.state_dict().keys()) == list(ours.state_dict().keys())
        assert [n for n, _ in ref.named_parameters()] == [n for n, _ in ours.named_parameters()]
        for (k, a), (_, b) in zip(ref.state_dict().items(), ours.state_dict().items()):
            assert torch.equal(a, b), k
        ours.load_state_dict(ref.state_dict())


def test_bf16_point_oracle_is_close_to_fp32_oracle():
    torch.manual_seed(0)
    g = O.init_weights(O.OracleResnet2D(3, 3, 2))
    x = torch.rand(1, 3, 32, 32) * 2 - 1
    y, yb = g(x), O.forward_bf16_points(g, x)
    assert ((y - yb).norm() / y.norm()).item() < 5e-2


def test_image_pool_matches_reference_semantics():
    random.seed(5)
    pool = O.OracleImagePool(2)
    a, b, c = torch.zeros(1, 1, 2, 2), torch.ones(1, 1, 2, 2), torch.full((1, 1, 2, 2), 2.0)
    assert torch.equal(pool.query(a), a) and torch.equal(pool.query(b), b)  # filling: returns the input
    out = pool.query(c)
    assert out.shape == c.shape
    from ganslate_b200.data.utils.image_pool import ImagePool
    random.seed(5)
    mine = ImagePool(2)
    mine.query(a), mine.query(b)
    assert torch.equal(mine.query(c), out)


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
def test_oracle_losses_equal_reference_classes():
    """PatchNCELoss, CycleGANLosses, Pix2PixLoss of the reference vs their oracle restatements."""
    from types import SimpleNamespace
    m = R.modules()
    conf = SimpleNamespace(train=SimpleNamespace(batch_size=2, gan=SimpleNamespace(optimizer=SimpleNamespace(
        nce_T=0.07, lambda_AB=10.0, lambda_BA=5.0, lambda_identity=0.5, proportion_ssim=0.0, lambda_pix2pix=30.0))))
    torch.manual_seed(0)
    q, k = torch.randn(2 * 64, 32), torch.randn(2 * 64, 32)
    assert torch.allclose(m["PatchNCELoss"](conf)(q, k), O.patchnce_loss(q, k, 2, 0.07), atol=1e-6)
    v = {n: torch.rand(2, 3, 8, 8) for n in ("real_A", "real_B", "fake_A", "fake_B", "rec_A", "rec_B", "idt_A", "idt_B")}
    ref = m["CycleGANLosses"](conf)(v)
    ora = O.cyclegan_losses(v, 10.0, 5.0, 0.5)
    assert set(ref) == set(ora)
    for n in ref:
        assert torch.allclose(ref[n], ora[n], atol=1e-6), n
    assert torch.allclose(m["Pix2PixLoss"](conf)(v["fake_B"], v["real_B"]), 30.0 * torch.nn.functional.l1_loss(v["fake_B"], v["real_B"]))


def test_b200_cut_modules_match_oracle_structure(monkeypatch):
    """FeaturePatchMLP state_dict keys / init order equal the oracle's (and hence the reference's cut.py:244-250); its
    forward (the fused gather + MLP + L2-norm entry point, here through the CPU restatement of the ABI) equals the
    oracle's module chain."""
    import fake_cabi
    fake_cabi.install(monkeypatch)
    from ganslate_b200.nn.gans.unpaired.cut import FeaturePatchMLP
    from ganslate_b200.nn.utils import init_weights
    torch.manual_seed(3)
    a = O.init_weights(O.OracleFeaturePatchMLP([3, 128, 256], 256, 64))
    torch.manual_seed(3)
    b = FeaturePatchMLP([3, 128, 256], 256, 64)
    init_weights(b, "normal", 0.02)
    assert list(a.state_dict().keys()) == list(b.state_dict().keys())
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), k
    feats = [torch.rand(2, c, 8, 8) for c in (3, 128, 256)]
    ids = [torch.arange(0, 64, 2)] * 3
    fa, _ = a(feats, ids)
    fb, _ = b(feats, ids)
    for x, y in zip(fa, fb):
        assert torch.allclose(x, y, atol=1e-6)


# ------------------------------------------------------------------------------------------------ 3-D oracle
def _digest_small(t):
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, 5).long()
    return {"sum": t.sum().item(), "abs_sum": t.abs().sum().item(), "samples": t[idx].tolist(), "numel": t.numel()}


def _vnet_small(use_inverse=True):
    from oracle import torch_oracle3d as O3
    net = O3.OracleVnet3D(1, 1, first_layer_channels=8, down_blocks=(1, 1), up_blocks=(1, 1), use_inverse=use_inverse)
    torch.manual_seed(0)  # seeded AFTER construction: the draws of init_weights do not depend on constructor RNG use
    return O.init_weights(net)


def test_oracle3d_matches_golden():
    """tests/golden/vnet3d_patchgan3d_small.json was generated from the REFERENCE's Vnet3D / PatchGAN3D
    (oracle/make_golden.py, over the memcnn stand-in): outputs and input gradients of both directions."""
    from oracle import torch_oracle3d as O3
    with open(os.path.join(GOLDEN, "vnet3d_patchgan3d_small.json")) as f:
        gold = json.load(f)
    g = _vnet_small()
    assert list(g.state_dict().keys()) == gold["vnet_keys"]
    d = O3.OraclePatchGAN3D(1, 16, 2, (4, 4, 4))
    torch.manual_seed(0)
    O.init_weights(d)
    assert list(d.state_dict().keys()) == gold["patchgan_keys"]
    gen = torch.Generator().manual_seed(3)
    x = (torch.rand((1, 1, 8, 16, 16), generator=gen) * 2 - 1).requires_grad_(True)
    for inverse in (False, True):
        y = g(x, inverse=inverse)
        (gx,) = torch.autograd.grad(y.square().sum(), x)
        assert_digest(_digest_small(y) | {"sq_sum": (y.double()**2).sum().item()}, gold["vnet"][str(inverse)]["y"],
                      f"vnet y inverse={inverse}")
        assert_digest(_digest_small(gx) | {"sq_sum": (gx.double()**2).sum().item()}, gold["vnet"][str(inverse)]["dx"],
                      f"vnet dx inverse={inverse}", rtol=1e-3)
    xd = torch.rand((1, 1, 16, 16, 16), generator=gen) * 2 - 1
    p = d(xd)
    assert_digest(_digest_small(p) | {"sq_sum": (p.double()**2).sum().item()}, gold["patchgan"]["y"], "patchgan y")


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
def test_oracle3d_networks_equal_reference_modules():
    from oracle import torch_oracle3d as O3
    ref = R.modules()
    torch.manual_seed(0)
    rg = ref["Vnet3D"](1, 1, "instance", first_layer_channels=8, down_blocks=(1, 1), up_blocks=(1, 1),
                       use_memory_saving=False, use_inverse=True)
    og = _vnet_small()
    assert list(og.state_dict().keys()) == list(rg.state_dict().keys())
    rg.load_state_dict(og.state_dict())
    x, _ = O3.synthetic_volume(1, 1, 8, 16, seed=3)
    for inverse in (False, True):
        assert torch.allclose(og(x, inverse=inverse), rg(x, inverse=inverse), atol=1e-6)
    rd = ref["PatchGAN3D"](1, 16, 2, (4, 4, 4), "instance")
    od = O3.OraclePatchGAN3D(1, 16, 2, (4, 4, 4))
    torch.manual_seed(0)
    O.init_weights(od)
    assert list(od.state_dict().keys()) == list(rd.state_dict().keys())
    rd.load_state_dict(od.state_dict())
    xd, _ = O3.synthetic_volume(1, 1, 16, 16, seed=4)
    assert torch.allclose(od(xd), rd(xd), atol=1e-6)


def test_oracle_piresnet3d_and_separable_vnet_match_golden():
    """tests/golden/piresnet3d_separable_small.json: the REFERENCE's Piresnet3D (piresnet3d.py:28-119) and
    Vnet3D(is_separable=True) (separable.py:5-83), generated by oracle/make_golden.py."""
    from oracle import torch_oracle3d as O3
    with open(os.path.join(GOLDEN, "piresnet3d_separable_small.json")) as f:
        gold = json.load(f)
    p = O3.OraclePiresnet3D(2, 2, 2, first_layer_channels=8, use_inverse=True)
    torch.manual_seed(0)
    O.init_weights(p)
    assert list(p.state_dict().keys()) == gold["piresnet"]["keys"]
    gen = torch.Generator().manual_seed(5)
    x = (torch.rand((1, 2, 8, 12, 12), generator=gen) * 2 - 1).requires_grad_(True)
    for inverse in (False, True):
        y = p(x, inverse=inverse)
        (gx,) = torch.autograd.grad(y.square().sum(), x)
        assert_digest(_digest_small(y) | {"sq_sum": (y.double()**2).sum().item()}, gold["piresnet"][str(inverse)]["y"],
                      f"piresnet y inverse={inverse}")
        assert_digest(_digest_small(gx) | {"sq_sum": (gx.double()**2).sum().item()}, gold["piresnet"][str(inverse)]["dx"],
                      f"piresnet dx inverse={inverse}", rtol=1e-3)
    r3 = O3.OracleResnet3D(1, 2, n_residual_blocks=1)
    torch.manual_seed(0)
    O.init_weights(r3)
    assert list(r3.state_dict().keys()) == gold["resnet3d"]["keys"]
    gen = torch.Generator().manual_seed(8)
    x = (torch.rand((1, 1, 8, 8, 8), generator=gen) * 2 - 1).requires_grad_(True)
    y = r3(x)
    (gx,) = torch.autograd.grad(y.square().sum(), x)
    assert_digest(_digest_small(y) | {"sq_sum": (y.double()**2).sum().item()}, gold["resnet3d"]["y"], "resnet3d y")
    assert_digest(_digest_small(gx) | {"sq_sum": (gx.double()**2).sum().item()}, gold["resnet3d"]["dx"], "resnet3d dx",
                  rtol=1e-3)
    v = O3.OracleVnet3D(1, 1, first_layer_channels=8, down_blocks=(1, 1), up_blocks=(1, 1), use_inverse=True,
                        is_separable=True)
    torch.manual_seed(0)
    O.init_weights(v)
    assert list(v.state_dict().keys()) == gold["vnet_separable"]["keys"]
    gen = torch.Generator().manual_seed(6)
    x = (torch.rand((1, 1, 8, 16, 16), generator=gen) * 2 - 1).requires_grad_(True)
    for inverse in (False, True):
        y = v(x, inverse=inverse)
        (gx,) = torch.autograd.grad(y.square().sum(), x)
        assert_digest(_digest_small(y) | {"sq_sum": (y.double()**2).sum().item()},
                      gold["vnet_separable"][str(inverse)]["y"], f"separable vnet y inverse={inverse}")
        assert_digest(_digest_small(gx) | {"sq_sum": (gx.double()**2).sum().item()},
                      gold["vnet_separable"][str(inverse)]["dx"], f"separable vnet dx inverse={inverse}", rtol=1e-3)


def test_b200_piresnet3d_and_separable_vnet_mirror_reference_state_dict():
    """Drop-in obligation: same state_dict keys, parameter order and shapes as the reference modules."""
    from ganslate_b200.nn.generators import Piresnet3D, Resnet3D, Vnet3D
    from oracle import torch_oracle3d as O3
    with open(os.path.join(GOLDEN, "piresnet3d_separable_small.json")) as f:
        gold = json.load(f)
    ours = Resnet3D(1, 2, "instance", n_residual_blocks=1)
    ref = O3.OracleResnet3D(1, 2, n_residual_blocks=1)
    assert list(ours.state_dict().keys()) == gold["resnet3d"]["keys"]
    assert [tuple(p.shape) for p in ours.parameters()] == [tuple(p.shape) for p in ref.parameters()]
    ours.load_state_dict(ref.state_dict())
    ours = Piresnet3D(2, 2, "instance", depth=2, first_layer_channels=8, use_memory_saving=False, use_inverse=True)
    ref = O3.OraclePiresnet3D(2, 2, 2, first_layer_channels=8, use_inverse=True)
    assert list(ours.state_dict().keys()) == gold["piresnet"]["keys"]
    assert [tuple(p.shape) for p in ours.parameters()] == [tuple(p.shape) for p in ref.parameters()]
    ours.load_state_dict(ref.state_dict())
    ours = Vnet3D(1, 1, "instance", first_layer_channels=8, down_blocks=(1, 1), up_blocks=(1, 1), use_memory_saving=False,
                  use_inverse=True, is_separable=True)
    ref = O3.OracleVnet3D(1, 1, first_layer_channels=8, down_blocks=(1, 1), up_blocks=(1, 1), use_inverse=True,
                          is_separable=True)
    assert list(ours.state_dict().keys()) == gold["vnet_separable"]["keys"]
    assert [tuple(p.shape) for p in ours.parameters()] == [tuple(p.shape) for p in ref.parameters()]
    ours.load_state_dict(ref.state_dict())
    # same seed -> same initial weights as the reference's init_weights (class names containing "Conv", module order)
    from ganslate_b200.nn.utils import init_weights
    torch.manual_seed(0)
    init_weights(ours)
    torch.manual_seed(0)
    O.init_weights(ref)
    for (k, a), (_, b) in zip(ours.state_dict().items(), ref.state_dict().items()):
        assert torch.equal(a, b), k


def test_oracle3d_coupling_is_invertible_and_revgan_step_runs():
    """The memcnn boundary has no reference golden values: pin it by invertibility of the shared core and by a
    complete RevGAN iteration producing finite losses and gradients for every generator parameter used."""
    from oracle import torch_oracle3d as O3
    g = _vnet_small()
    core = g.downs[0].core
    h = torch.randn(1, 16, 4, 8, 8)
    assert torch.allclose(core(core(h), inverse=True), h, atol=1e-5)
    random.seed(0)
    m = O3.OracleRevGAN(O3.default_3d_conf(first_layer_channels=8, down_blocks=(1, 1), up_blocks=(1, 1), ndf=16,
                                           n_layers=2), seed=0)
    a, b = O3.synthetic_volume(1, 1, 16, 16, seed=1)
    losses = m.optimize_parameters(a, b, step_optimizers=False)
    assert set(losses) == {"G_AB", "G_BA", "cycle_A", "cycle_B", "D_B", "D_A"}
    assert all(torch.isfinite(torch.tensor(v)) for v in losses.values())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.networks["G"].parameters())


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("num_downs,use_dropout", [(5, False), (7, True)])
def test_oracle_unet2d_equals_reference_module(num_downs, use_dropout):
    """OracleUnet2D (oracle/torch_oracle.py) against ganslate/nn/generators/unet/unet2d.py: same keys, same-seed
    weights, identical outputs and gradients (eval mode when the blocks carry Dropout), and the B200 module mirrors
    the state_dict and the initialisation."""
    from ganslate_b200.nn.generators import Unet2D
    from ganslate_b200.nn.utils import init_weights as b200_init
    m = R.modules()
    torch.manual_seed(5)
    ref = m["Unet2D"](3, 2, num_downs, "instance", ngf=8, use_dropout=use_dropout)
    m["init_weights"](ref, "normal", 0.02)
    torch.manual_seed(5)
    ora = O.init_weights(O.OracleUnet2D(3, 2, num_downs, ngf=8, use_dropout=use_dropout))
    torch.manual_seed(5)
    ours = Unet2D(3, 2, num_downs, "instance", ngf=8, use_dropout=use_dropout)
    b200_init(ours, "normal", 0.02)
    for other in (ora, ours):
        assert list(ref.state_dict().keys()) == list(other.state_dict().keys())
        for (k, a), (_, b) in zip(ref.state_dict().items(), other.state_dict().items()):
            assert torch.equal(a, b), k
    ref.eval(), ora.eval()
    size = 2 ** num_downs
    x = (torch.rand(2, 3, size, size) * 2 - 1).requires_grad_(True)
    xo = x.detach().clone().requires_grad_(True)
    yr, yo = ref(x), ora(xo)
    assert yr.shape == (2, 2, size, size) and torch.allclose(yr, yo, atol=1e-6)
    g = torch.randn_like(yr)
    yr.backward(g), yo.backward(g)
    assert torch.allclose(x.grad, xo.grad, atol=1e-6)
    for (k, a), (_, b) in zip(ref.named_parameters(), ora.named_parameters()):
        assert torch.allclose(a.grad, b.grad, atol=1e-5), k


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("shape", [(2, 3, 40, 37), (1, 2, 4, 24, 30)])
def test_oracle_ssim_equals_reference(shape):
    """oracle ssim_distance / cycle_loss against ganslate/nn/losses/utils/ssim.py and CycleLoss (values and gradients)."""
    R.setup()
    from ganslate.nn.losses.utils.ssim import SSIMLoss
    from ganslate.nn.losses.cyclegan_losses import CycleLoss
    torch.manual_seed(0)
    real = torch.rand(shape) * 2 - 1
    rec = (0.8 * real + 0.2 * (torch.rand(shape) * 2 - 1))
    a, b = rec.clone().requires_grad_(True), rec.clone().requires_grad_(True)
    lr = SSIMLoss()((a + 1) / 2, (real + 1) / 2, data_range=1)
    lo = O.ssim_distance((b + 1) / 2, (real + 1) / 2, 1.0)
    assert torch.allclose(lr, lo, atol=1e-7)
    lr.backward(), lo.backward()
    assert torch.allclose(a.grad, b.grad, atol=1e-8)
    a.grad = b.grad = None
    cr = CycleLoss(0.84)(real, a)
    co = O.cycle_loss(real, b, 0.84)
    assert torch.allclose(cr, co, atol=1e-7)
    cr.backward(), co.backward()
    assert torch.allclose(a.grad, b.grad, atol=1e-8)


def test_oracle_unet2d_and_ssim_match_reference_golden():
    """tests/golden/unet2d_ssim.json was produced by the reference's Unet2D / SSIMLoss / CycleLoss
    (oracle/make_golden.py::reference_unet2d_ssim); this check needs no reference tree."""
    with open(os.path.join(GOLDEN, "unet2d_ssim.json")) as f:
        gold = json.load(f)
    u = gold["unet2d"]
    c = u["config"]
    torch.manual_seed(c["seed"])
    net = O.init_weights(O.OracleUnet2D(c["in_channels"], c["out_channels"], c["num_downs"], ngf=c["ngf"]))
    assert list(net.state_dict().keys()) == u["keys"]
    gen = torch.Generator().manual_seed(c["data_seed"])
    x = (torch.rand(tuple(c["shape"]), generator=gen) * 2 - 1).requires_grad_(True)
    y = net(x)
    y.square().sum().backward()
    assert_digest(digest(y), u["y"], "unet y")
    assert_digest(digest(x.grad), u["dx"], "unet dx", rtol=1e-3)
    for k, p in net.named_parameters():
        assert_digest(digest(p.grad), u["grads"][k], k, rtol=1e-3)
    for case in gold["ssim"]:
        gen = torch.Generator().manual_seed(case["data_seed"])
        real = torch.rand(tuple(case["shape"]), generator=gen) * 2 - 1
        rec = (0.8 * real + 0.2 * (torch.rand(tuple(case["shape"]), generator=gen) * 2 - 1)).requires_grad_(True)
        l = O.ssim_distance((rec + 1) / 2, (real + 1) / 2, 1.0)
        (g,) = torch.autograd.grad(l, rec)
        assert close(float(l), case["ssim"], 1e-5), (float(l), case["ssim"])
        assert_digest(digest(g), case["d_ssim"], "d ssim", rtol=1e-3, atol=1e-9)
        lc = O.cycle_loss(real, rec, 0.84)
        (gc,) = torch.autograd.grad(lc, rec)
        assert close(float(lc), case["cycle_084"], 1e-5)
        assert_digest(digest(gc), case["d_cycle_084"], "d cycle", rtol=1e-3, atol=1e-9)


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
def test_oracle_unet3d_equals_reference_module():
    from ganslate_b200.nn.generators import Unet3D
    from ganslate_b200.nn.utils import init_weights as b200_init
    R.setup()
    from ganslate.nn.generators.unet.unet3d import Unet3D as RefUnet3D
    m = R.modules()
    torch.manual_seed(6)
    ref = RefUnet3D(1, 1, 5, "instance", ngf=8)
    m["init_weights"](ref, "normal", 0.02)
    torch.manual_seed(6)
    ora = O.init_weights(O.OracleUnet3D(1, 1, 5, ngf=8))
    torch.manual_seed(6)
    ours = Unet3D(1, 1, 5, "instance", ngf=8)
    b200_init(ours, "normal", 0.02)
    for other in (ora, ours):
        assert list(ref.state_dict().keys()) == list(other.state_dict().keys())
        for (k, a), (_, b) in zip(ref.state_dict().items(), other.state_dict().items()):
            assert torch.equal(a, b), k
    x = (torch.rand(1, 1, 32, 32, 32) * 2 - 1).requires_grad_(True)
    xo = x.detach().clone().requires_grad_(True)
    yr, yo = ref(x), ora(xo)
    assert torch.allclose(yr, yo, atol=1e-6)
    g = torch.randn_like(yr)
    yr.backward(g), yo.backward(g)
    assert torch.allclose(x.grad, xo.grad, atol=1e-6)
