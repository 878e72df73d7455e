"""Pix2Pix reconstruction loss -- API of ganslate/nn/losses/pix2pix_losses.py:8-19 (lambda * L1(fake_B, real_B))."""
from ganslate_b200 import ops


class Pix2PixLoss:

    def __init__(self, conf):
        self.lambda_pix2pix = conf.train.gan.optimizer.lambda_pix2pix

    def __call__(self, fake_B, real_B):
        return self.lambda_pix2pix * ops.L1Fn.apply(fake_B, real_B)
