"""Cycle-consistency and identity losses -- API of ganslate/nn/losses/cyclegan_losses.py:7-101.
L1 terms run the fused value+gradient kernel (gb_l1), the SSIM term the stencil kernels (gb_ssim_fwd / _bwd)."""
from ganslate_b200 import ops


class CycleGANLosses:

    def __init__(self, conf):
        opt = conf.train.gan.optimizer
        self.lambda_AB = opt.lambda_AB
        self.lambda_BA = opt.lambda_BA
        self.criterion_cycle = CycleLoss(opt.proportion_ssim)
        self.criterion_idt = IdentityLoss(opt.lambda_identity) if opt.lambda_identity > 0 else None

    def is_using_identity(self):
        return self.criterion_idt is not None

    def __call__(self, visuals):
        losses = {
            'cycle_A': self.lambda_AB * self.criterion_cycle(visuals['real_A'], visuals['rec_A']),
            'cycle_B': self.lambda_BA * self.criterion_cycle(visuals['real_B'], visuals['rec_B']),
        }
        if self.criterion_idt:
            if visuals['idt_A'] is None or visuals['idt_B'] is None:
                raise ValueError("idt_A and/or idt_B is not computed but the identity loss is defined.")
            losses['idt_B'] = self.lambda_AB * self.criterion_idt(visuals['idt_B'], visuals['real_B'])
            losses['idt_A'] = self.lambda_BA * self.criterion_idt(visuals['idt_A'], visuals['real_A'])
        return losses


class CycleLoss:
    """cyclegan_losses.py:60-91: L1, or alpha * SSIM-distance((x + 1) / 2) + (1 - alpha) * L1 when proportion_ssim > 0
    (the SSIM stencil runs in csrc/ssim.cu)."""

    def __init__(self, proportion_ssim):
        self.alpha = float(proportion_ssim)
        self.beta = 1.0 - self.alpha

    def __call__(self, real, reconstructed):
        cycle_loss_l1 = ops.L1Fn.apply(reconstructed, real)
        if self.alpha > 0:
            # (x + 1) / 2 folded into the kernel's input mapping; data_range 1 (cyclegan_losses.py:84-88)
            cycle_loss_ssim = ops.SsimFn.apply(reconstructed, real, 0.5, 0.5, 1.0)
            return self.alpha * cycle_loss_ssim + self.beta * cycle_loss_l1
        return cycle_loss_l1


class IdentityLoss:

    def __init__(self, lambda_identity):
        self.lambda_identity = lambda_identity

    def __call__(self, idt, real):
        return ops.L1Fn.apply(idt, real) * self.lambda_identity
