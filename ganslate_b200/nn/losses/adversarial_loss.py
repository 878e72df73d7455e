"""GAN objectives -- API of ganslate/nn/losses/adversarial_loss.py:7-98.

`lsgan` (the reference default, configs/base.py:21) runs the fused sm_100a reduction kernel that produces the
loss and its gradient in one pass; `vanilla` / `wgangp` are tiny elementwise ATen reductions over the
30x30 patch map."""
from typing import Dict, Union

import torch
from torch import nn

from ganslate_b200 import ops


class AdversarialLoss(nn.Module):

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0):
        super().__init__()
        self.register_buffer('real_label', torch.tensor(target_real_label))
        self.register_buffer('fake_label', torch.tensor(target_fake_label))
        self._real, self._fake = float(target_real_label), float(target_fake_label)
        self.gan_mode = gan_mode
        if gan_mode == 'vanilla':
            self.loss = nn.BCEWithLogitsLoss()
        elif gan_mode in ('lsgan', 'wgangp'):
            self.loss = None
        else:
            raise NotImplementedError(f"GAN mode {gan_mode} not implemented.")

    def get_target_tensor(self, prediction, target_is_real):
        target = self.real_label if target_is_real else self.fake_label
        return target.expand_as(prediction)

    def calculate_loss(self, prediction: torch.Tensor, target_is_real: bool):
        if self.gan_mode == 'lsgan':
            return ops.MseConstFn.apply(prediction, self._real if target_is_real else self._fake)
        if self.gan_mode == 'vanilla':
            return self.loss(prediction, self.get_target_tensor(prediction, target_is_real))
        return -prediction.mean() if target_is_real else prediction.mean()

    def forward(self, prediction: Union[Dict[str, torch.Tensor], torch.Tensor], target_is_real: bool):
        if isinstance(prediction, dict):
            losses = [self.calculate_loss(pred, target_is_real) for pred in prediction.values()]
            return torch.stack(losses).mean()
        return self.calculate_loss(prediction, target_is_real)
