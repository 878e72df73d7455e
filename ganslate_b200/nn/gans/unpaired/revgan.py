"""RevGAN recipe -- one training iteration as ganslate/nn/gans/unpaired/revgan.py:89-212 orders it: ONE partially
invertible generator (Vnet3D with `use_inverse`) serves both directions (`inverse=True` is B -> A), the G step runs
with frozen discriminators, then D_B and D_A train on (real, pooled fake).

Kept reference behaviours: backward_G feeds D_B with fake_A and D_A with fake_B (revgan.py:196-197), backward_D
calls backward with retain_graph=True (revgan.py:183-186)."""
import itertools
from dataclasses import dataclass, field

import torch

from ganslate_b200 import configs
from ganslate_b200.data.utils.image_pool import make_image_pool
from ganslate_b200.nn.gans.base import BaseGAN
from ganslate_b200.nn.gans.unpaired import cyclegan
from ganslate_b200.nn.losses.adversarial_loss import AdversarialLoss
from ganslate_b200.nn.losses.cyclegan_losses import CycleGANLosses


@dataclass
class OptimizerConfig(cyclegan.OptimizerConfig):
    pass


@dataclass
class RevGANConfig(configs.base.BaseGANConfig):
    pool_size: int = 50
    optimizer: OptimizerConfig = field(default_factory=OptimizerConfig)


class RevGAN(BaseGAN):

    def __init__(self, conf):
        super().__init__(conf)
        self.visuals = {n: None for n in ['real_A', 'fake_B', 'rec_A', 'idt_A', 'real_B', 'fake_A', 'rec_B', 'idt_B']}
        self.losses = {n: None for n in ['G_AB', 'D_B', 'cycle_A', 'idt_A', 'G_BA', 'D_A', 'cycle_B', 'idt_B']}
        self.optimizers = {'G': None, 'D': None}
        self.networks = {n: None for n in (['G', 'D_B', 'D_A'] if self.is_train else ['G'])}  # revgan.py:50
        if self.is_train:
            self.fake_A_pool = make_image_pool(conf)
            self.fake_B_pool = make_image_pool(conf)
        self.setup()

    def init_criterions(self):
        self.criterion_adv = AdversarialLoss(self.conf.train.gan.optimizer.adversarial_loss_type).to(self.device)
        self.criterion_G = CycleGANLosses(self.conf)

    def init_optimizers(self):
        o = self.conf.train.gan.optimizer
        params_D = itertools.chain(self.networks['D_B'].parameters(), self.networks['D_A'].parameters())
        self.optimizers['G'] = self.make_adam(self.networks['G'].parameters(), o.lr_G, (o.beta1, o.beta2))
        self.optimizers['D'] = self.make_adam(params_D, o.lr_D, (o.beta1, o.beta2))

    def set_input(self, input):
        self.visuals['real_A'] = self.stage_input('real_A', input['A'])
        self.visuals['real_B'] = self.stage_input('real_B', input['B'])

    def optimize_parameters(self):
        """One iteration in the reference's order (revgan.py:89-116).  With `train.cuda_graph` (one GPU; a data-parallel
        run falls back to eager DistributedDataParallel, BaseGAN.parallelize_networks) the generator phase and the
        discriminator phase are captured once and replayed; the image pools stay host logic in between, as in CycleGAN."""
        if self.graph_mode('step'):
            self.run_graphed('G', self._phase_G)
            pool_query = cyclegan.CycleGAN._pool_query
            fake_B = self.stage_input('pool_B', pool_query(self.fake_B_pool, self.visuals['fake_B']))
            fake_A = self.stage_input('pool_A', pool_query(self.fake_A_pool, self.visuals['fake_A']))
            self.run_graphed('D', lambda: self._phase_D(fake_B, fake_A))
            return
        with self.eager_stream():
            self._phase_G()
            self._phase_D(None, None)

    def _phase_G(self):
        discriminators = [self.networks['D_B'], self.networks['D_A']]
        self.forward()
        self.metrics.update(self.training_metrics.compute_metrics_G(self.visuals))
        self.set_requires_grad(discriminators, False)
        self.optimizers['G'].zero_grad(set_to_none=True)
        self.backward_G()
        self.optimizers['G'].step()

    def _phase_D(self, fake_B, fake_A):
        """fake_B / fake_A: pooled fakes already staged (graph replay), or None: query the pools here."""
        self.set_requires_grad([self.networks['D_B'], self.networks['D_A']], True)
        self.optimizers['D'].zero_grad(set_to_none=True)
        self.backward_D('D_B', fake_B)
        self.metrics.update(self.training_metrics.compute_metrics_D('D_B', self.pred_real, self.pred_fake))
        self.backward_D('D_A', fake_A)
        self.metrics.update(self.training_metrics.compute_metrics_D('D_A', self.pred_real, self.pred_fake))
        self.optimizers['D'].step()

    def forward(self):
        G = self.networks['G']
        real_A, real_B = self.visuals['real_A'], self.visuals['real_B']
        if self._streams() is not None:
            self._prepack(['G'])

        # the two cycles are independent chains through the one generator (train.multi_stream: two CUDA streams)
        def chain_A():
            fake_B = G(real_A)
            return fake_B, G(fake_B, inverse=True)

        def chain_B():
            fake_A = G(real_B, inverse=True)
            return fake_A, G(fake_A)

        (fake_B, rec_A), (fake_A, rec_B) = self._fork_join(chain_A, chain_B)
        idt_B, idt_A = None, None
        if self.criterion_G.is_using_identity():
            idt_B, idt_A = self._fork_join(lambda: G(real_B), lambda: G(real_A, inverse=True))
        self.visuals.update({'fake_B': fake_B, 'rec_A': rec_A, 'idt_A': idt_A, 'fake_A': fake_A, 'rec_B': rec_B,
                             'idt_B': idt_B})

    def backward_D(self, discriminator, pooled_fake=None):
        if discriminator == 'D_B':
            real = self.visuals['real_B']
            fake = pooled_fake if pooled_fake is not None else self.fake_B_pool.query(self.visuals['fake_B'])
        elif discriminator == 'D_A':
            real = self.visuals['real_A']
            fake = pooled_fake if pooled_fake is not None else self.fake_A_pool.query(self.visuals['fake_A'])
        else:
            raise ValueError('The discriminator has to be either "D_A" or "D_B".')
        self.pred_real = self.networks[discriminator](real)
        self.pred_fake = self.networks[discriminator](fake.detach())
        loss_real = self.criterion_adv(self.pred_real, target_is_real=True)
        loss_fake = self.criterion_adv(self.pred_fake, target_is_real=False)
        self.losses[discriminator] = loss_real + loss_fake
        self.backward(loss=self.losses[discriminator], optimizer=self.optimizers['D'], retain_graph=True,
                      loss_id=0 if discriminator == 'D_B' else 1)

    def backward_G(self):
        if self._streams() is not None:
            self._prepack(['D_B', 'D_A'])
        # the reference's pairing (revgan.py:196-197): D_B sees fake_A, D_A sees fake_B
        pred_B, pred_A = self._fork_join(lambda: self.networks['D_B'](self.visuals['fake_A']),
                                         lambda: self.networks['D_A'](self.visuals['fake_B']))
        self.losses['G_AB'] = self.criterion_adv(pred_B, target_is_real=True)
        self.losses['G_BA'] = self.criterion_adv(pred_A, target_is_real=True)
        losses_G = self.criterion_G(self.visuals)
        self.losses.update(losses_G)
        combined_loss_G = sum(losses_G.values()) + self.losses['G_AB'] + self.losses['G_BA']
        self.backward(loss=combined_loss_G, optimizer=self.optimizers['G'], loss_id=2)

    def infer(self, input, direction='AB'):
        assert direction in ['AB', 'BA'], "Specify which generator direction, AB or BA, to use."
        with torch.no_grad():
            return self.networks['G'](input, inverse=(direction == 'BA'))
