"""Pix2Pix (conditional GAN) recipe -- one iteration as ganslate/nn/gans/paired/pix2pix.py:76-152 orders it:
G(A), G step on D(cat[A, G(A)]) + lambda * L1, then D on (cat[A, B], cat[A, G(A).detach()])."""
from dataclasses import dataclass, field

import torch

from ganslate_b200 import configs
from ganslate_b200.nn.gans.base import BaseGAN
from ganslate_b200.nn.losses.adversarial_loss import AdversarialLoss
from ganslate_b200.nn.losses.pix2pix_losses import Pix2PixLoss


@dataclass
class OptimizerConfig(configs.base.BaseOptimizerConfig):
    lambda_pix2pix: float = 100.0


@dataclass
class Pix2PixConditionalGANConfig(configs.base.BaseGANConfig):
    optimizer: OptimizerConfig = field(default_factory=OptimizerConfig)


class Pix2PixConditionalGAN(BaseGAN):

    def __init__(self, conf):
        super().__init__(conf)
        self.visuals = {n: None for n in ['real_A', 'fake_B', 'real_B']}
        self.losses = {n: None for n in ['G', 'D', 'pix2pix']}
        self.optimizers = {'G': None, 'D': None}
        self.networks = {n: None for n in (['G', 'D'] if self.is_train else ['G'])}
        self.setup()

    def init_criterions(self):
        self.criterion_adv = AdversarialLoss(self.conf.train.gan.optimizer.adversarial_loss_type).to(self.device)
        self.criterion_pix2pix = Pix2PixLoss(self.conf)

    def init_optimizers(self):
        o = self.conf.train.gan.optimizer
        self.optimizers['G'] = self.make_adam(self.networks['G'].parameters(), o.lr_G, (o.beta1, o.beta2))
        self.optimizers['D'] = self.make_adam(self.networks['D'].parameters(), o.lr_D, (o.beta1, o.beta2))

    def set_input(self, input):
        self.visuals['real_A'] = self.stage_input('real_A', input['A'])
        self.visuals['real_B'] = self.stage_input('real_B', input['B'])

    graph_sync = True  # optimize_parameters issues the flat-bucket all-reduces itself (BaseGAN.parallelize_networks)

    def optimize_parameters(self):
        """One iteration in the reference's order (pix2pix.py:76-101).  Data parallel + CUDA graphs: DDP's hooks are not
        captured, so the gradients are averaged explicitly between the segments, as in CycleGAN: the discriminator
        phase reads neither the generator's weights nor anything produced after `forward()`, so it runs while the
        generator bucket is being reduced and both optimizers step afterwards -- same result as the reference order."""
        sync = self.grad_syncs
        if sync is None:
            if self.graph_mode('step'):
                self.run_graphed('step', self._step)
                return
            with self.eager_stream():
                self._step()
            return
        if self.graph_mode('step'):
            self.run_graphed('G', self._phase_G)
            sync['G'].launch()
            self.run_graphed('D', self._phase_D)
            sync['D'].launch()
            sync['G'].finish()
            self.run_graphed('stepG', self.optimizers['G'].step)
            sync['D'].finish()
            self.run_graphed('stepD', self.optimizers['D'].step)
            return
        with self.eager_stream():
            self._phase_G()
            sync['G'].launch()
            self._phase_D()
            sync['D'].launch()
            sync['G'].finish()
            self.optimizers['G'].step()
            sync['D'].finish()
            self.optimizers['D'].step()

    def _phase_G(self):
        self.forward()
        self.metrics.update(self.training_metrics.compute_metrics_G(self.visuals))
        # ---- G (D frozen: its weight-gradient kernels are skipped)
        self.set_requires_grad(self.networks['D'], False)
        self.optimizers['G'].zero_grad(set_to_none=True)
        self.backward_G()

    def _phase_D(self):
        self.set_requires_grad(self.networks['D'], True)
        self.optimizers['D'].zero_grad(set_to_none=True)
        self.backward_D()
        self.metrics.update(self.training_metrics.compute_metrics_D('D', self.pred_real, self.pred_fake))

    def _step(self):
        self._phase_G()
        self.optimizers['G'].step()
        self._phase_D()
        self.optimizers['D'].step()

    def backward_G(self):
        real_A, real_B, fake_B = self.visuals['real_A'], self.visuals['real_B'], self.visuals['fake_B']
        pred = self.networks['D'](torch.cat([real_A, fake_B], dim=1))  # D(A, G(A))
        self.losses['G'] = self.criterion_adv(pred, target_is_real=True)
        self.losses['pix2pix'] = self.criterion_pix2pix(fake_B, real_B)
        self.backward(loss=self.losses['G'] + self.losses['pix2pix'], optimizer=self.optimizers['G'])

    def backward_D(self):
        real_A, real_B, fake_B = self.visuals['real_A'], self.visuals['real_B'], self.visuals['fake_B']
        D = self.networks['D']
        if self._streams("real-fake") is not None:
            self._prepack(['D'])
        # the real and the fake pass are independent (train.multi_stream: two CUDA streams, forward and backward)
        self.pred_real, self.pred_fake = self._fork_join(lambda: D(torch.cat([real_A, real_B], dim=1)),
                                                         lambda: D(torch.cat([real_A, fake_B.detach()], dim=1)),
                                                         tag="real-fake")
        loss_real = self.criterion_adv(self.pred_real, target_is_real=True)
        loss_fake = self.criterion_adv(self.pred_fake, target_is_real=False)
        self.losses['D'] = loss_real + loss_fake
        self.backward(loss=self.losses['D'], optimizer=self.optimizers['D'])

    def forward(self):
        self.visuals.update({'fake_B': self.networks['G'](self.visuals['real_A'])})

    def infer(self, input):
        with torch.no_grad():
            return self.networks['G'].forward(input)
