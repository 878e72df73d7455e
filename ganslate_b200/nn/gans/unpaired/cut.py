"""CUT / FastCUT recipe -- one iteration as ganslate/nn/gans/unpaired/cut.py:113-226 orders it: D step first, then
G + patch-MLP on adversarial + PatchNCE (+ identity PatchNCE) losses.  The features come from the generator's
`encoder` layer list through the fused kernels (`layers.run_encoder`); the reference's `.encoder` access breaks under
DDP, here the wrapper is unwrapped first so CUT also runs data-parallel."""
from dataclasses import dataclass, field
from typing import Tuple

import numpy as np
import torch
from torch import nn
from torch.nn.parallel import DistributedDataParallel

from ganslate_b200 import configs, ops
from ganslate_b200.nn import layers
from ganslate_b200.nn.gans.base import BaseGAN
from ganslate_b200.nn.losses.adversarial_loss import AdversarialLoss
from ganslate_b200.nn.losses.cut_losses import PatchNCELoss
from ganslate_b200.nn.utils import get_network_device, init_net


@dataclass
class OptimizerConfig(configs.base.BaseOptimizerConfig):
    lambda_adv: float = 1
    lambda_nce: float = 1
    lambda_nce_idt: float = 0.5
    nce_T: float = 0.07


@dataclass
class CUTConfig(configs.base.BaseGANConfig):
    nce_layers: Tuple[int] = (0, 4, 8, 12, 16)
    mlp_nc: int = 256
    num_patches: int = 256
    use_equivariance_flip: bool = False
    optimizer: OptimizerConfig = field(default_factory=OptimizerConfig)


def _unwrap(net):
    return net.module if isinstance(net, DistributedDataParallel) else net


class CUT(BaseGAN):

    def __init__(self, conf):
        super().__init__(conf)
        o = conf.train.gan.optimizer
        self.lambda_adv, self.lambda_nce, self.lambda_nce_idt = o.lambda_adv, o.lambda_nce, o.lambda_nce_idt
        self.nce_layers = tuple(conf.train.gan.nce_layers)
        self.num_patches = conf.train.gan.num_patches
        self.use_equivariance_flip = conf.train.gan.use_equivariance_flip
        self.is_flipped = False
        self.visuals = {n: None for n in ['real_A', 'fake_B', 'real_B', 'idt_B']}
        self.losses = {n: None for n in ['D', 'G', 'NCE', 'NCE_idt']}
        self.networks = {n: None for n in (['G', 'D', 'mlp'] if self.is_train else ['G'])}
        self.fixed_patch_ids = None  # tests inject the reference's random patch ids here
        self.setup()

    def init_networks(self):
        super().init_networks()
        if self.is_train:
            in_ch = self.conf.train.gan.generator.in_out_channels.AB[0]
            channels = probe_network_channels(self.networks['G'], self.nce_layers, in_ch)
            mlp = FeaturePatchMLP(channels, self.conf.train.gan.num_patches, self.conf.train.gan.mlp_nc)
            self.networks['mlp'] = init_net(mlp, self.conf, self.device)

    def init_optimizers(self):
        o = self.conf.train.gan.optimizer
        betas = (o.beta1, o.beta2)
        self.optimizers['G'] = self.make_adam(self.networks['G'].parameters(), o.lr_G, betas)
        self.optimizers['D'] = self.make_adam(self.networks['D'].parameters(), o.lr_D, betas)
        self.optimizers['mlp'] = self.make_adam(self.networks['mlp'].parameters(), o.lr_G, betas)

    def init_criterions(self):
        self.criterion_adv = AdversarialLoss(self.conf.train.gan.optimizer.adversarial_loss_type).to(self.device)
        self.criterion_nce = [PatchNCELoss(self.conf).to(self.device) for _ in self.nce_layers]

    def save_checkpoint(self, iter_idx):
        super().save_checkpoint(iter_idx)  # the reference drops the mlp optimizer as well (base.py:244-245)

    graph_sync = True  # optimize_parameters issues the flat-bucket all-reduces itself (BaseGAN.parallelize_networks)

    def optimize_parameters(self):
        """One iteration in the reference's order (cut.py:113-137).  With `train.cuda_graph` the two phases (forward + D
        step, G + patch-MLP step) are captured once and replayed; the discriminator is stepped BEFORE the second phase
        (the generator's adversarial loss goes through the updated discriminator).  The patch ids are drawn on the
        device inside the captured region (torch.randperm under the graph-safe Philox generator), so every replay
        samples new patches.  Data parallel + graphs: explicit flat-bucket all-reduces between the segments, as in
        CycleGAN.  (The capture path has not run on a B200 yet: bench.py keeps CUT eager unless --graph is given.)"""
        sync = self.grad_syncs
        if self.graph_mode('step'):
            if self.use_equivariance_flip:
                raise RuntimeError("train.cuda_graph cannot replay use_equivariance_flip (a host-side coin flip per "
                                   "iteration changes the captured program)")
            self.run_graphed('D', lambda: self._phase_D(step=sync is None))
            if sync:
                sync['D'].launch()
                sync['D'].finish()
                self.run_graphed('stepD', self.optimizers['D'].step)
            self.run_graphed('G', lambda: self._phase_G(step=sync is None))
            if sync:
                sync['G'].launch()
                sync['mlp'].launch()
                sync['G'].finish()
                self.run_graphed('stepG', self.optimizers['G'].step)
                sync['mlp'].finish()
                self.run_graphed('stepMLP', self.optimizers['mlp'].step)
            return
        with self.eager_stream():
            self._phase_D(step=sync is None)
            if sync:
                sync['D'].launch()
                sync['D'].finish()
                self.optimizers['D'].step()
            self._phase_G(step=sync is None)
            if sync:
                sync['G'].launch()
                sync['mlp'].launch()
                sync['G'].finish()
                self.optimizers['G'].step()
                sync['mlp'].finish()
                self.optimizers['mlp'].step()

    def _phase_D(self, step=True):
        self.forward()
        self.set_requires_grad(self.networks['D'], True)
        self.optimizers['D'].zero_grad(set_to_none=True)
        self.backward_D()
        if step:
            self.optimizers['D'].step()

    def _phase_G(self, step=True):
        self.set_requires_grad(self.networks['D'], False)
        self.optimizers['G'].zero_grad(set_to_none=True)
        self.optimizers['mlp'].zero_grad(set_to_none=True)
        self.backward_G_and_mlp()
        if step:
            self.optimizers['G'].step()
            self.optimizers['mlp'].step()

    def set_input(self, input):
        self.visuals['real_A'] = self.stage_input('real_A', input['A'])
        self.visuals['real_B'] = self.stage_input('real_B', input['B'])

    def forward(self):
        using_idt = self.lambda_nce_idt > 0
        real_A = self.visuals['real_A']
        real_B = self.visuals['real_B'] if using_idt else None
        if self.use_equivariance_flip and self.is_train:
            self.is_flipped = np.random.random() > 0.5
            if self.is_flipped:
                real_A = real_A.flip(-1)
                if using_idt:
                    real_B = real_B.flip(-1)
        G = self.networks['G']
        if not using_idt:
            self.visuals['fake_B'] = G(real_A)
            return
        # the translation and the identity pass are independent (train.multi_stream: two CUDA streams, forward and
        # backward -- at CUT's batch 1 a single chain leaves most SMs idle)
        if self._streams() is not None:
            self._prepack(['G'])
        self.visuals['fake_B'], self.visuals['idt_B'] = self._fork_join(lambda: G(real_A), lambda: G(real_B))

    def backward_D(self):
        D = self.networks['D']
        if self._streams("real-fake") is not None:
            self._prepack(['D'])
        pred_real, pred_fake = self._fork_join(lambda: D(self.visuals['real_B']),
                                               lambda: D(self.visuals['fake_B'].detach()), tag="real-fake")
        loss_real = self.criterion_adv(pred_real, True).mean()
        loss_fake = self.criterion_adv(pred_fake, False).mean()
        self.losses['D'] = loss_real + loss_fake
        self.backward(loss=self.losses['D'], optimizer=self.optimizers['D'], loss_id=0)

    def backward_G_and_mlp(self):
        real_A, real_B = self.visuals['real_A'], self.visuals['real_B']
        fake_B, idt_B = self.visuals['fake_B'], self.visuals['idt_B']
        adversarial_loss = 0
        if self.lambda_adv > 0:
            pred_fake = self.networks['D'](fake_B)
            adversarial_loss = self.criterion_adv(pred_fake, True).mean() * self.lambda_adv
            self.losses['G'] = adversarial_loss
        nce_loss = 0
        if self.lambda_nce > 0:
            if self._streams("nce") is not None:
                self._prepack(['G'])
            if self.lambda_nce_idt > 0:
                # (the two contrastive terms are independent: two CUDA streams with train.multi_stream)
                nce_loss, nce_idt = self._fork_join(lambda: self._calculate_nce_loss(real_A, fake_B),
                                                    lambda: self._calculate_nce_loss(real_B, idt_B), tag="nce")
            else:
                nce_loss = self._calculate_nce_loss(real_A, fake_B)
            self.losses['NCE'] = nce_loss
            if self.lambda_nce_idt > 0:
                nce_idt_loss = self.lambda_nce_idt * nce_idt
                nce_loss = (1 - self.lambda_nce_idt) * nce_loss + nce_idt_loss
                self.losses['NCE_idt'] = nce_idt_loss
        self.backward(loss=adversarial_loss + nce_loss, optimizer=(self.optimizers['G'], self.optimizers['mlp']), loss_id=1)

    def _calculate_nce_loss(self, source, target):
        G = _unwrap(self.networks['G'])
        source_feats, target_feats = self._fork_join(lambda: extract_features(source, G, self.nce_layers),
                                                     lambda: extract_features(target, G, self.nce_layers), tag="feats")
        if self.is_flipped:
            target_feats = [feat.flip(-1) for feat in target_feats]
        source_pool, patch_ids = self.networks['mlp'](source_feats, self.fixed_patch_ids)
        target_pool, _ = self.networks['mlp'](target_feats, patch_ids)
        nce_loss = 0
        for target_feat, source_feat, criterion in zip(target_pool, source_pool, self.criterion_nce):
            loss = criterion(target_feat, source_feat) * self.lambda_nce
            nce_loss = nce_loss + loss.mean()
        return nce_loss / len(self.nce_layers)


class FeaturePatchMLP(nn.Module):
    """cut.py:229-282: gather `num_patches` positions (same ids for every batch item), 2-layer MLP, L2 normalise.
    The nn.Linear modules are parameter containers (same state_dict keys `mlps.{i}.{0,2}.{weight,bias}` and init as
    the reference); gather + both layers + the normalisation run as ONE fused sm_100a launch per feature
    (ops.PatchMlpFn -> csrc/patch_mlp.cu), fp32 FMAs, no library GEMM."""

    def __init__(self, channels_per_feature, num_patches=256, nc=256):
        super().__init__()
        self.num_patches = num_patches
        self.l2norm = LNorm(2)
        self.mlps = nn.ModuleList(
            [nn.Sequential(nn.Linear(c, nc), nn.ReLU(), nn.Linear(nc, nc)) for c in channels_per_feature])

    def forward(self, feats, patch_ids=None):
        device = feats[0].device
        return_feats, return_ids = [], []
        for i, feat in enumerate(feats):
            n_pos = feat[0, 0].numel()
            if self.num_patches > 0:
                if patch_ids is not None:
                    patch_id = patch_ids[i]
                else:
                    patch_id = torch.randperm(n_pos, device=device)
                    patch_id = patch_id[:int(min(self.num_patches, len(patch_id)))]
                ids = patch_id
            else:
                patch_id, ids = [], torch.arange(n_pos, device=device)
            lin1, lin2 = self.mlps[i][0], self.mlps[i][2]
            feat_patch = ops.PatchMlpFn.apply(feat, ids, lin1.weight, lin1.bias, lin2.weight, lin2.bias)
            return_feats.append(feat_patch)
            return_ids.append(patch_id)
        return return_feats, return_ids


class LNorm(nn.Module):

    def __init__(self, power=2):
        super().__init__()
        self.power = power

    def forward(self, x):
        norm = x.pow(self.power).sum(1, keepdim=True).pow(1. / self.power)
        return x.div(norm + 1e-7)


def extract_features(input, network, layers_to_extract_from):
    """Features after the listed indices of `network.encoder` (cut.py:297-312), computed by the fused kernels."""
    assert len(network.encoder) >= max(layers_to_extract_from), \
        f"The encoder has {len(network.encoder)} layers, cannot extract features from layers that do not exist."
    return layers.run_encoder(network, list(network.encoder), input, layers_to_extract_from)


def probe_network_channels(network, layers_of_interest, input_channels=3):
    """Channel count of every tapped feature (cut.py:315-333), from one small dry run."""
    device = get_network_device(network)
    with torch.no_grad():
        shape = (1, input_channels, 16, 64, 64) if '3d' in str(network).lower() else (1, input_channels, 64, 64)
        feats = extract_features(torch.zeros(shape, device=device), _unwrap(network), tuple(layers_of_interest))
    return [f.shape[1] for f in feats]
