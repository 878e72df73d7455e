"""SSIM distance loss kernels (csrc/ssim.cu) against the CPU oracle (oracle/torch_oracle.py::ssim_distance, pinned to
ganslate/nn/losses/utils/ssim.py in tests/test_oracle.py).  fp32 stencil arithmetic: loss within 1e-4 relative,
gradient within 2e-3 of its own maximum (the gradient divides by sqrt(S), which amplifies rounding where the two
images nearly agree)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(2, 3, 64, 64), (1, 1, 37, 53), (3, 2, 11, 75), (1, 2, 4, 32, 45), (1, 3, 256, 256)])
def test_ssim_loss_and_gradient_vs_oracle(shape):
    from ganslate_b200 import ops
    from oracle import torch_oracle as O
    torch.manual_seed(1)
    real = torch.rand(shape) * 2 - 1
    rec = (0.7 * real + 0.3 * (torch.rand(shape) * 2 - 1))
    r_cpu = rec.clone().requires_grad_(True)
    lo = O.ssim_distance((r_cpu + 1) / 2, (real + 1) / 2, 1.0)
    (3.0 * lo).backward()
    r_gpu = rec.clone().cuda().requires_grad_(True)
    lg = ops.SsimFn.apply(r_gpu, real.cuda(), 0.5, 0.5, 1.0)
    (3.0 * lg).backward()
    torch.cuda.synchronize()
    assert abs(lg.item() - lo.item()) <= 1e-4 * abs(lo.item()), (lg.item(), lo.item())
    err = (r_gpu.grad.cpu() - r_cpu.grad).abs().max().item()
    assert err <= 2e-3 * r_cpu.grad.abs().max().item(), (err, r_cpu.grad.abs().max().item())


def test_cycle_loss_with_ssim_matches_oracle():
    from types import SimpleNamespace as NS
    from ganslate_b200.nn.losses.cyclegan_losses import CycleGANLosses
    from oracle import torch_oracle as O
    torch.manual_seed(2)
    conf = NS(train=NS(gan=NS(optimizer=NS(lambda_AB=10.0, lambda_BA=5.0, lambda_identity=0.5, proportion_ssim=0.84))))
    crit = CycleGANLosses(conf)
    names = ("real_A", "real_B", "fake_A", "fake_B", "rec_A", "rec_B", "idt_A", "idt_B")
    cpu = {k: (torch.rand(2, 3, 48, 48) * 2 - 1) for k in names}
    for k in ("rec_A", "rec_B", "idt_A", "idt_B"):
        cpu[k].requires_grad_(True)
    gpu = {k: v.detach().clone().cuda().requires_grad_(v.requires_grad) for k, v in cpu.items()}
    lo = O.cyclegan_losses(cpu, 10.0, 5.0, 0.5, proportion_ssim=0.84)
    lg = crit(gpu)
    sum(lo.values()).backward()
    sum(lg.values()).backward()
    torch.cuda.synchronize()
    assert set(lo) == set(lg)
    for k in lo:
        assert abs(lg[k].item() - lo[k].item()) <= 1e-4 * abs(lo[k].item()), k
    for k in ("rec_A", "rec_B", "idt_A", "idt_B"):
        err = (gpu[k].grad.cpu() - cpu[k].grad).abs().max().item()
        assert err <= 2e-3 * cpu[k].grad.abs().max().item(), (k, err)


def test_training_ssim_metric_matches_reference_formula():
    """train.metrics.ssim (ganslate/utils/metrics/train_metrics.py:36-47,56-67): ssim_A = 1 - SSIMLoss((real_A + 1) / 2,
    (rec_A + 1) / 2, data_range=1), no gradient; here the stencil kernel with the input mapping folded in."""
    from types import SimpleNamespace as NS
    from ganslate_b200.nn.gans.base import TrainingMetricsLite
    from oracle import torch_oracle as O

    class Conf(dict):
        def get(self, k, d=None):
            return dict.get(self, k, d)
    conf = NS(train=Conf(metrics=Conf(ssim=True, discriminator_evolution=False)))
    tm = TrainingMetricsLite(conf)
    g = torch.Generator().manual_seed(3)
    real = torch.rand(2, 3, 48, 40, generator=g) * 2 - 1
    rec = (real + 0.2 * torch.randn(2, 3, 48, 40, generator=g)).clamp(-1, 1)
    vis = {"real_A": real.cuda(), "rec_A": rec.cuda().requires_grad_(True), "real_B": rec.cuda(), "rec_B": real.cuda()}
    m = tm.compute_metrics_G(vis)
    ref_a = 1 - O.ssim_distance((real + 1) / 2, (rec + 1) / 2, 1.0)
    ref_b = 1 - O.ssim_distance((rec + 1) / 2, (real + 1) / 2, 1.0)
    assert abs(float(m["ssim_A"]) - float(ref_a)) <= 1e-4 * abs(float(ref_a)) + 1e-6
    assert abs(float(m["ssim_B"]) - float(ref_b)) <= 1e-4 * abs(float(ref_b)) + 1e-6
    assert not m["ssim_A"].requires_grad
    assert TrainingMetricsLite(NS(train=Conf(metrics=Conf(ssim=False)))).compute_metrics_G(vis) == {}
