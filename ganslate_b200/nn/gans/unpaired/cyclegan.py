"""CycleGAN recipe -- one training iteration exactly as ganslate/nn/gans/unpaired/cyclegan.py:92-214 orders it:
4 generator passes, G step with frozen discriminators, then D_B and D_A on (real, pooled fake)."""
import itertools
from dataclasses import dataclass, field

import torch

from ganslate_b200 import configs
from ganslate_b200.data.utils.image_pool import make_image_pool
from ganslate_b200.nn.gans.base import BaseGAN
from ganslate_b200.nn.losses.adversarial_loss import AdversarialLoss
from ganslate_b200.nn.losses.cyclegan_losses import CycleGANLosses



@dataclass
class OptimizerConfig(configs.base.BaseOptimizerConfig):
    lambda_AB: float = 10.0
    lambda_BA: float = 10.0
    lambda_identity: float = 0
    proportion_ssim: float = 0.84


@dataclass
class CycleGANConfig(configs.base.BaseGANConfig):
    pool_size: int = 50
    optimizer: OptimizerConfig = field(default_factory=OptimizerConfig)


class CycleGAN(BaseGAN):

    def __init__(self, conf):
        super().__init__(conf)
        self.visuals = {n: None for n in ['real_A', 'fake_B', 'rec_A', 'idt_A', 'real_B', 'fake_A', 'rec_B', 'idt_B']}
        self.losses = {n: None for n in ['G_AB', 'D_B', 'cycle_A', 'idt_A', 'G_BA', 'D_A', 'cycle_B', 'idt_B']}
        self.optimizers = {'G': None, 'D': None}
        # dict order = construction and weight-init order (cyclegan.py:52; base.py:51-67)
        self.networks = {n: None for n in (['G_AB', 'G_BA', 'D_B', 'D_A'] if self.is_train else ['G_AB'])}
        if self.is_train:
            self.fake_A_pool = make_image_pool(conf)
            self.fake_B_pool = make_image_pool(conf)
        self.setup()

    def init_criterions(self):
        self.criterion_adv = AdversarialLoss(self.conf.train.gan.optimizer.adversarial_loss_type).to(self.device)
        self.criterion_G = CycleGANLosses(self.conf)

    def init_optimizers(self):
        o = self.conf.train.gan.optimizer
        params_G = itertools.chain(self.networks['G_AB'].parameters(), self.networks['G_BA'].parameters())
        params_D = itertools.chain(self.networks['D_B'].parameters(), self.networks['D_A'].parameters())
        self.optimizers['G'] = self.make_adam(params_G, o.lr_G, (o.beta1, o.beta2))
        self.optimizers['D'] = self.make_adam(params_D, o.lr_D, (o.beta1, o.beta2))

    def set_input(self, input):
        self.visuals['real_A'] = self.stage_input('real_A', input['A'])
        self.visuals['real_B'] = self.stage_input('real_B', input['B'])

    graph_sync = True  # optimize_parameters issues the flat-bucket all-reduces itself (BaseGAN.parallelize_networks)

    def optimize_parameters(self):
        """One iteration in the reference's order (cyclegan.py:92-124).  With `train.cuda_graph` the two halves
        (forward + G step, D steps) are captured once and replayed; the ImagePool stays host logic in between."""
        sync = self.grad_syncs  # data parallel + graphs: explicit all-reduce between the graph segments
        if self.graph_mode('step'):
            self.run_graphed('G', lambda: self._phase_G(step=sync is None))
            if sync:
                sync['G'].launch()  # NCCL on a side stream, overlaps with the discriminator graph
            fake_B = self.stage_input('pool_B', self._pool_query(self.fake_B_pool, self.visuals['fake_B']))
            fake_A = self.stage_input('pool_A', self._pool_query(self.fake_A_pool, self.visuals['fake_A']))
            self.run_graphed('D', lambda: self._phase_D(fake_B, fake_A, step=sync is None))
            if sync:
                # same result as the reference order (G step before the D phase): the D phase reads neither the
                # generators' weights nor anything produced after `forward()`
                sync['D'].launch()
                sync['G'].finish()
                self.run_graphed('stepG', self.optimizers['G'].step)
                sync['D'].finish()
                self.run_graphed('stepD', self.optimizers['D'].step)
            return
        with self.eager_stream():
            self._phase_G(step=sync is None)
            if sync:
                sync['G'].launch()
            self._phase_D(None, None, step=sync is None)
            if sync:
                sync['D'].launch()
                sync['G'].finish()
                self.optimizers['G'].step()
                sync['D'].finish()
                self.optimizers['D'].step()

    @staticmethod
    def _pool_query(pool, images):
        """Graph mode: the pool must not keep a reference into the graph's static output buffers.  The device pool
        copies the images into its own storage; the reference's list-based pool stores the tensors it is given."""
        from ganslate_b200.data.utils.image_pool import DeviceImagePool
        images = images.detach()
        return pool.query(images if isinstance(pool, DeviceImagePool) else images.clone())

    def _phase_G(self, step=True):
        discriminators = [self.networks['D_B'], self.networks['D_A']]
        self.forward()
        self.metrics.update(self.training_metrics.compute_metrics_G(self.visuals))
        # ---- G_AB and G_BA (discriminators frozen: their weight-gradient kernels are skipped)
        self.set_requires_grad(discriminators, False)
        self.optimizers['G'].zero_grad(set_to_none=True)
        self.backward_G()
        if step:
            self.optimizers['G'].step()

    def _phase_D(self, fake_B, fake_A, step=True):
        discriminators = [self.networks['D_B'], self.networks['D_A']]
        self.set_requires_grad(discriminators, True)
        self.optimizers['D'].zero_grad(set_to_none=True)
        if self._streams() is not None:
            self._prepack(['D_B', 'D_A'])

            def one(name, fake):
                self.backward_D(name, fake)
                return self.training_metrics.compute_metrics_D(name, self.pred_real, self.pred_fake)

            m_B, m_A = self._fork_join(lambda: one('D_B', fake_B), lambda: one('D_A', fake_A))
            self.metrics.update(m_B)
            self.metrics.update(m_A)
        else:
            self.backward_D('D_B', fake_B)
            self.metrics.update(self.training_metrics.compute_metrics_D('D_B', self.pred_real, self.pred_fake))
            self.backward_D('D_A', fake_A)
            self.metrics.update(self.training_metrics.compute_metrics_D('D_A', self.pred_real, self.pred_fake))
        if step:
            self.optimizers['D'].step()

    def forward(self):
        real_A, real_B = self.visuals['real_A'], self.visuals['real_B']
        G_AB, G_BA = self.networks['G_AB'], self.networks['G_BA']
        if self._streams() is not None:
            self._prepack(['G_AB', 'G_BA'])

        def chain_A():
            fake_B = G_AB(real_A)
            return fake_B, G_BA(fake_B)

        def chain_B():
            fake_A = G_BA(real_B)
            return fake_A, G_AB(fake_A)

        (fake_B, rec_A), (fake_A, rec_B) = self._fork_join(chain_A, chain_B)
        idt_B, idt_A = None, None
        if self.criterion_G.is_using_identity():
            idt_B, idt_A = self._fork_join(lambda: G_AB(real_B), lambda: G_BA(real_A))
        self.visuals.update({'fake_B': fake_B, 'rec_A': rec_A, 'idt_A': idt_A, 'fake_A': fake_A, 'rec_B': rec_B,
                             'idt_B': idt_B})

    def backward_D(self, discriminator, fake=None):
        """`fake`: pooled fake images when the caller already queried the ImagePool (graph mode)."""
        if discriminator == 'D_B':
            real = self.visuals['real_B']
            fake = self.fake_B_pool.query(self.visuals['fake_B']) if fake is None else fake
        elif discriminator == 'D_A':
            real = self.visuals['real_A']
            fake = self.fake_A_pool.query(self.visuals['fake_A']) if fake is None else fake
        else:
            raise ValueError('The discriminator has to be either "D_A" or "D_B".')
        D = self.networks[discriminator]
        self.pred_real, self.pred_fake = self._fork_join(lambda: D(real), lambda: D(fake.detach()), tag="real-fake")
        loss_real = self.criterion_adv(self.pred_real, target_is_real=True)
        loss_fake = self.criterion_adv(self.pred_fake, target_is_real=False)
        self.losses[discriminator] = loss_real + loss_fake
        self.backward(loss=self.losses[discriminator], optimizer=self.optimizers['D'], loss_id=2)

    def backward_G(self):
        if self._streams() is not None:
            self._prepack(['D_B', 'D_A'])
        # (their own stream pair: in the backward pass a discriminator's data gradient then runs beside the second
        #  generator of the same chain instead of in front of it)
        pred_B, pred_A = self._fork_join(lambda: self.networks['D_B'](self.visuals['fake_B']),
                                         lambda: self.networks['D_A'](self.visuals['fake_A']), tag="D")
        self.losses['G_AB'] = self.criterion_adv(pred_B, target_is_real=True)
        self.losses['G_BA'] = self.criterion_adv(pred_A, target_is_real=True)
        losses_G = self.criterion_G(self.visuals)
        self.losses.update(losses_G)
        combined_loss_G = sum(losses_G.values()) + self.losses['G_AB'] + self.losses['G_BA']
        self.backward(loss=combined_loss_G, optimizer=self.optimizers['G'], loss_id=0)

    def infer(self, input, direction='AB'):
        assert direction in ['AB', 'BA'], "Specify which generator direction, AB or BA, to use."
        with torch.no_grad():
            return self.networks[f'G_{direction}'](input)
