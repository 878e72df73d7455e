"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.json from the REFERENCE's own modules.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
The reference has no golden vectors of its own; these fixtures are produced by importing
its nn modules (oracle/reference_import.py) and driving them through one CycleGAN iteration exactly as
ganslate/nn/gans/unpaired/cyclegan.py:92-214 does (the recipe classes themselves do not import on py3.12).
"""
import itertools
import json
import os
import random
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_import as R  # noqa: E402


def tensor_digest(t):
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, 5).long()
    return {"sum": t.sum().item(), "abs_sum": t.abs().sum().item(), "sq_sum": (t * t).sum().item(),
            "samples": t[idx].tolist(), "numel": t.numel()}


def reference_cyclegan_step(size, n_blocks, seed=0, data_seed=1, batch=1, lambda_identity=0.0):
    m = R.modules()
    torch.manual_seed(seed)
    random.seed(0)
    nets = {}
    for name in ("G_AB", "G_BA", "D_B", "D_A"):  # cyclegan.py:52 order
        net = m["Resnet2D"](3, 3, "instance", n_blocks) if name[0] == "G" else m["PatchGAN2D"](3, 64, 3, (4, 4), "instance")
        m["init_weights"](net, "normal", 0.02)
        nets[name] = net
    conf = SimpleNamespace(train=SimpleNamespace(gan=SimpleNamespace(optimizer=SimpleNamespace(
        lambda_AB=10.0, lambda_BA=10.0, lambda_identity=lambda_identity, proportion_ssim=0.0))))
    crit_adv = m["AdversarialLoss"]("lsgan")
    crit_G = m["CycleGANLosses"](conf)
    opt_G = torch.optim.Adam(itertools.chain(nets["G_AB"].parameters(), nets["G_BA"].parameters()), lr=2e-4, betas=(0.5, 0.999))
    opt_D = torch.optim.Adam(itertools.chain(nets["D_B"].parameters(), nets["D_A"].parameters()), lr=2e-4, betas=(0.5, 0.999))
    pool_A, pool_B = m["ImagePool"](50), m["ImagePool"](50)
    g = torch.Generator().manual_seed(data_seed)
    real_A = torch.rand((batch, 3, size, size), generator=g) * 2 - 1
    real_B = torch.rand((batch, 3, size, size), generator=g) * 2 - 1

    def set_rg(ns, flag):
        for n in ns:
            for p in n.parameters():
                p.requires_grad = flag

    v = {"real_A": real_A, "real_B": real_B}
    v["fake_B"] = nets["G_AB"](real_A)
    v["rec_A"] = nets["G_BA"](v["fake_B"])
    v["fake_A"] = nets["G_BA"](real_B)
    v["rec_B"] = nets["G_AB"](v["fake_A"])
    v["idt_A"] = v["idt_B"] = None
    if crit_G.is_using_identity():
        v["idt_B"] = nets["G_AB"](real_B)
        v["idt_A"] = nets["G_BA"](real_A)
    losses = {}
    set_rg([nets["D_B"], nets["D_A"]], False)
    opt_G.zero_grad(set_to_none=True)
    losses["G_AB"] = crit_adv(nets["D_B"](v["fake_B"]), target_is_real=True)
    losses["G_BA"] = crit_adv(nets["D_A"](v["fake_A"]), target_is_real=True)
    lg = crit_G(v)
    losses.update(lg)
    (sum(lg.values()) + losses["G_AB"] + losses["G_BA"]).backward()
    grads = {f"{n}.{k}": tensor_digest(p.grad) for n in ("G_AB", "G_BA") for k, p in nets[n].named_parameters()}
    opt_G.step()
    set_rg([nets["D_B"], nets["D_A"]], True)
    opt_D.zero_grad(set_to_none=True)
    for name, real, fake, pool in (("D_B", real_B, v["fake_B"], pool_B), ("D_A", real_A, v["fake_A"], pool_A)):
        fake = pool.query(fake)
        pr, pf = nets[name](real), nets[name](fake.detach())
        losses[name] = crit_adv(pr, target_is_real=True) + crit_adv(pf, target_is_real=False)
        losses[name].backward()
    grads.update({f"{n}.{k}": tensor_digest(p.grad) for n in ("D_B", "D_A") for k, p in nets[n].named_parameters()})
    opt_D.step()
    weights = {f"{n}.{k}": tensor_digest(p) for n in nets for k, p in nets[n].named_parameters()}
    return {
        "config": {"size": size, "n_blocks": n_blocks, "seed": seed, "data_seed": data_seed, "batch": batch,
                   "lambda_identity": lambda_identity},
        "losses": {k: float(x) for k, x in losses.items()},
        "visuals": {k: tensor_digest(v[k]) for k in ("fake_B", "rec_A", "fake_A", "rec_B")},
        "grads": grads,
        "weights_after_step": weights,
        "state_dict_keys": {n: list(nets[n].state_dict().keys()) for n in nets},
        "param_counts": {n: sum(p.numel() for p in nets[n].parameters()) for n in nets},
    }


def reference_3d_small():
    """Reference Vnet3D (both directions, over the memcnn stand-in) and PatchGAN3D on a 1x1x8x16x16 volume (PatchGAN3D: 1x1x16x16x16):
    outputs and input gradients of sum(y^2)."""
    m = R.modules()
    g = m["Vnet3D"](1, 1, "instance", first_layer_channels=8, down_blocks=(1, 1), up_blocks=(1, 1),
                    use_memory_saving=False, use_inverse=True)
    torch.manual_seed(0)
    m["init_weights"](g, "normal", 0.02)
    d = m["PatchGAN3D"](1, 16, 2, (4, 4, 4), "instance")
    torch.manual_seed(0)
    m["init_weights"](d, "normal", 0.02)
    gen = torch.Generator().manual_seed(3)
    x = (torch.rand((1, 1, 8, 16, 16), generator=gen) * 2 - 1).requires_grad_(True)
    out = {"vnet_keys": list(g.state_dict().keys()), "patchgan_keys": list(d.state_dict().keys()), "vnet": {}}
    for inverse in (False, True):
        y = g(x, inverse=inverse)
        (gx,) = torch.autograd.grad(y.square().sum(), x)
        out["vnet"][str(inverse)] = {"y": tensor_digest(y), "dx": tensor_digest(gx)}
    xd = torch.rand((1, 1, 16, 16, 16), generator=gen) * 2 - 1
    out["patchgan"] = {"y": tensor_digest(d(xd))}
    return out


def reference_piresnet_separable():
    """Reference Piresnet3D (piresnet3d.py, both directions) and Vnet3D(is_separable=True) (separable.py) over the
    memcnn stand-in: outputs and input gradients of sum(y^2) on small volumes."""
    R.setup()
    from ganslate.nn.generators.resnet.piresnet3d import Piresnet3D
    m = R.modules()
    out = {}
    p = Piresnet3D(2, 2, "instance", depth=2, first_layer_channels=8, use_memory_saving=False, use_inverse=True)
    torch.manual_seed(0)
    m["init_weights"](p, "normal", 0.02)
    gen = torch.Generator().manual_seed(5)
    x = (torch.rand((1, 2, 8, 12, 12), generator=gen) * 2 - 1).requires_grad_(True)
    out["piresnet"] = {"config": dict(in_channels=2, out_channels=2, depth=2, first_layer_channels=8, seed=0, data_seed=5,
                                      shape=[1, 2, 8, 12, 12]), "keys": list(p.state_dict().keys())}
    for inverse in (False, True):
        y = p(x, inverse=inverse)
        (gx,) = torch.autograd.grad(y.square().sum(), x)
        out["piresnet"][str(inverse)] = {"y": tensor_digest(y), "dx": tensor_digest(gx)}
    from ganslate.nn.generators.resnet.resnet3d import Resnet3D
    r3 = Resnet3D(1, 2, "instance", n_residual_blocks=1)
    torch.manual_seed(0)
    m["init_weights"](r3, "normal", 0.02)
    gen = torch.Generator().manual_seed(8)
    x = (torch.rand((1, 1, 8, 8, 8), generator=gen) * 2 - 1).requires_grad_(True)
    y = r3(x)
    (gx,) = torch.autograd.grad(y.square().sum(), x)
    out["resnet3d"] = {"config": dict(in_channels=1, out_channels=2, n_residual_blocks=1, seed=0, data_seed=8,
                                      shape=[1, 1, 8, 8, 8]), "keys": list(r3.state_dict().keys()),
                       "y": tensor_digest(y), "dx": tensor_digest(gx)}
    v = m["Vnet3D"](1, 1, "instance", first_layer_channels=8, down_blocks=(1, 1), up_blocks=(1, 1),
                    use_memory_saving=False, use_inverse=True, is_separable=True)
    torch.manual_seed(0)
    m["init_weights"](v, "normal", 0.02)
    gen = torch.Generator().manual_seed(6)
    x = (torch.rand((1, 1, 8, 16, 16), generator=gen) * 2 - 1).requires_grad_(True)
    out["vnet_separable"] = {"keys": list(v.state_dict().keys())}
    for inverse in (False, True):
        y = v(x, inverse=inverse)
        (gx,) = torch.autograd.grad(y.square().sum(), x)
        out["vnet_separable"][str(inverse)] = {"y": tensor_digest(y), "dx": tensor_digest(gx)}
    return out


def reference_unet2d_ssim():
    """Reference Unet2D (ganslate/nn/generators/unet/unet2d.py, dropout off) output / gradients on a fixed input, and
    reference SSIMLoss / CycleLoss(0.84) values and gradients (nn/losses/utils/ssim.py, cyclegan_losses.py:60-91)."""
    m = R.modules()
    from ganslate.nn.losses.utils.ssim import SSIMLoss
    from ganslate.nn.losses.cyclegan_losses import CycleLoss
    out = {}
    torch.manual_seed(5)
    net = m["Unet2D"](3, 2, 5, "instance", ngf=8, use_dropout=False)
    m["init_weights"](net, "normal", 0.02)
    gen = torch.Generator().manual_seed(7)
    x = (torch.rand((2, 3, 32, 64), generator=gen) * 2 - 1).requires_grad_(True)
    y = net(x)
    y.square().sum().backward()
    out["unet2d"] = {"config": dict(in_channels=3, out_channels=2, num_downs=5, ngf=8, seed=5, data_seed=7, shape=[2, 3, 32, 64]),
                     "keys": list(net.state_dict().keys()), "y": tensor_digest(y), "dx": tensor_digest(x.grad),
                     "grads": {k: tensor_digest(p.grad) for k, p in net.named_parameters()}}
    out["ssim"] = []
    for shape in ([2, 3, 40, 37], [1, 2, 4, 24, 30]):
        gen = torch.Generator().manual_seed(11)
        real = torch.rand(shape, generator=gen) * 2 - 1
        rec = (0.8 * real + 0.2 * (torch.rand(shape, generator=gen) * 2 - 1)).requires_grad_(True)
        l_ssim = SSIMLoss()((rec + 1) / 2, (real + 1) / 2, data_range=1)
        (g_ssim,) = torch.autograd.grad(l_ssim, rec)
        l_cyc = CycleLoss(0.84)(real, rec)
        (g_cyc,) = torch.autograd.grad(l_cyc, rec)
        out["ssim"].append({"shape": shape, "data_seed": 11, "ssim": float(l_ssim), "d_ssim": tensor_digest(g_ssim),
                            "cycle_084": float(l_cyc), "d_cycle_084": tensor_digest(g_cyc)})
    return out


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "unet2d_ssim.json"), "w") as f:
        json.dump(reference_unet2d_ssim(), f)
    print("unet2d_ssim written")
    with open(os.path.join(out_dir, "piresnet3d_separable_small.json"), "w") as f:
        json.dump(reference_piresnet_separable(), f)
    print("piresnet3d_separable_small written")
    with open(os.path.join(out_dir, "vnet3d_patchgan3d_small.json"), "w") as f:
        json.dump(reference_3d_small(), f)
    print("vnet3d_patchgan3d_small written")
    cases = {"cyclegan_step_32px_2blk": dict(size=32, n_blocks=2),
             "cyclegan_step_64px_3blk_idt": dict(size=64, n_blocks=3, lambda_identity=0.5, batch=2)}
    for name, kw in cases.items():
        rep = reference_cyclegan_step(**kw)
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            json.dump(rep, f)
        print(name, {k: round(v, 6) for k, v in rep["losses"].items()})


if __name__ == "__main__":
    main()
