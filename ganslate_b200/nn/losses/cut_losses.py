"""PatchNCE loss -- API of ganslate/nn/losses/cut_losses.py:5-43 on the fused sm_100a kernel (gb_patchnce_fwd/bwd)."""
from torch import nn

from ganslate_b200 import ops


class PatchNCELoss(nn.Module):

    def __init__(self, conf):
        super().__init__()
        self.batch_size = conf.train.batch_size
        self.nce_T = conf.train.gan.optimizer.nce_T

    def forward(self, feat_q, feat_k):
        """feat_q, feat_k: (batch * patches, dim). Returns the per-row loss (reduction 'none'); feat_k is treated
        as a constant exactly like `feat_k.detach()` in the reference (:16)."""
        return ops.PatchNCEFn.apply(feat_q, feat_k, self.batch_size, self.nce_T)
