"""Functional stand-in for memcnn (unpinned third-party dependency of the reference, absent from this image).

Restates the published algorithm of memcnn.AdditiveCoupling / memcnn.InvertibleModuleWrapper as called at
ganslate/nn/invertible.py:15-19:  y1 = x1 + Fm(x2); y2 = x2 + Gm(y1)  on a channel split in halves, inverse
x2 = y2 - Gm(y1); x1 = y1 - Fm(x2).  The real wrapper only changes WHEN activations are stored (recompute in
backward), not the values or gradients, so a plain autograd version gives the same numbers.  parity unpinned:
no reference test holds golden values for memcnn.
"""
import copy

import torch
from torch import nn


class AdditiveCoupling(nn.Module):

    def __init__(self, Fm, Gm=None, split_dim=1):
        super().__init__()
        self.Fm = Fm
        self.Gm = copy.deepcopy(Fm) if Gm is None else Gm
        self.split_dim = split_dim

    def forward(self, x):
        x1, x2 = torch.chunk(x, 2, dim=self.split_dim)
        y1 = x1 + self.Fm(x2)
        y2 = x2 + self.Gm(y1)
        return torch.cat([y1, y2], dim=self.split_dim)

    def inverse(self, y):
        y1, y2 = torch.chunk(y, 2, dim=self.split_dim)
        x2 = y2 - self.Gm(y1)
        x1 = y1 - self.Fm(x2)
        return torch.cat([x1, x2], dim=self.split_dim)


class InvertibleModuleWrapper(nn.Module):

    def __init__(self, fn, keep_input=False, keep_input_inverse=False, num_bwd_passes=1, disable=False,
                 preserve_rng_state=False):
        super().__init__()
        self._fn = fn
        self.keep_input = keep_input
        self.keep_input_inverse = keep_input_inverse
        self.disable = disable

    def forward(self, *x):
        return self._fn(*x)

    def inverse(self, *y):
        return self._fn.inverse(*y)
