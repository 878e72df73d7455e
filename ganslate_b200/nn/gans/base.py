"""GAN recipe base class -- the API of ganslate/nn/gans/base.py:16-321 (set_input / forward /
optimize_parameters / infer / checkpoints / DDP wrap / set_requires_grad) on top of the B200 modules.

Mixed precision: the reference's option is NVIDIA Apex AMP fp16 (base.py:118-126), off in every shipped config.
Here bf16 storage with fp32 accumulation IS the compute path of the kernels; the Apex switch is rejected."""
import logging
import os
from abc import ABC, abstractmethod
from pathlib import Path

import torch
from torch.nn.parallel import DistributedDataParallel

from ganslate_b200.nn.utils import get_scheduler
from ganslate_b200.utils import communication
from ganslate_b200.utils.builders import build_D, build_G


class TrainingMetricsLite:
    """The step-path part of ganslate/utils/metrics/train_metrics.py:10-67: mean discriminator outputs
    (`discriminator_evolution`) and the SSIM of the reconstructions (`ssim`: 1 - SSIMLoss((x + 1) / 2, (y + 1) / 2,
    data_range=1), train_metrics.py:36-47,56-67) -- the same stencil kernel the SSIM cycle loss runs (csrc/ssim.cu)."""

    def __init__(self, conf):
        m = conf.train.get("metrics", None)
        self.output_distributions = bool(m and m.get("discriminator_evolution", False))
        self.ssim = bool(m and m.get("ssim", False))

    def compute_metrics_D(self, name, pred_real, pred_fake):
        if not self.output_distributions:
            return {}
        if isinstance(pred_real, dict):
            pred_real, pred_fake = pred_real[next(iter(pred_real))], pred_fake[next(iter(pred_fake))]
        return {f"{name}_real": pred_real.detach().mean(), f"{name}_fake": pred_fake.detach().mean()}

    def get_SSIM_metric(self, input, target):
        from ganslate_b200 import ops
        with torch.no_grad():
            # (x + 1) / 2 is folded into the kernel's input mapping (in_scale 0.5, in_shift 0.5), data_range 1
            return 1 - ops.SsimFn.apply(input.detach(), target.detach(), 0.5, 0.5, 1.0)

    def compute_metrics_G(self, visuals):
        out = {}
        if not self.ssim:
            return out
        if visuals.get("rec_A") is not None and visuals.get("real_A") is not None:
            out["ssim_A"] = self.get_SSIM_metric(visuals["real_A"], visuals["rec_A"])
        if visuals.get("rec_B") is not None and visuals.get("real_B") is not None:
            out["ssim_B"] = self.get_SSIM_metric(visuals["real_B"], visuals["rec_B"])
        return out


class BaseGAN(ABC):
    # True for recipes whose optimize_parameters() drives `self.grad_syncs` (explicit gradient all-reduce between the
    # CUDA-graph segments); for the others data parallelism + train.cuda_graph falls back to eager DDP
    graph_sync = False

    def __init__(self, conf):
        self.logger = logging.getLogger("ganslate_b200")
        self.conf = conf
        self.is_train = self.conf.mode == "train"
        self.device = self._specify_device()
        self.output_dir = conf[conf.mode].output_dir
        self.visuals, self.metrics, self.losses, self.optimizers, self.networks = {}, {}, {}, {}, {}
        # CUDA-graph replay of the training step (launch-bound at batch 1: ~1.4 k kernels per iteration)
        self.use_cuda_graph = bool(conf[conf.mode].get("cuda_graph", False)) if self.is_train else False
        self.graph_warmup_iters = int(conf[conf.mode].get("cuda_graph_warmup", 11)) if self.is_train else 0
        self._graphs, self._static, self._graph_calls = {}, {}, 0
        self.grad_syncs = None  # {optimizer name: FlatGradSync} when graphs + data parallel
        self.input_copy_stream = None  # train.input_prefetch (graph mode): H2D of pinned inputs on its own stream
        if self.use_cuda_graph and bool(conf[conf.mode].get("input_prefetch", False)) and self.device.type == "cuda":
            self.input_copy_stream = torch.cuda.Stream(device=self.device)
        self._staging = {}
        self.graph_launches_per_step = 0

    def init_networks(self):
        """base.py:49-67: names starting with G/D, suffix _BA / _A selects direction / domain."""
        for name in self.networks.keys():
            if name.startswith('G'):
                self.networks[name] = build_G(self.conf, 'BA' if name.endswith('_BA') else 'AB', self.device)
            elif name.startswith('D'):
                self.networks[name] = build_D(self.conf, 'A' if name.endswith('_A') else 'B', self.device)

    @abstractmethod
    def init_criterions(self):
        """Initialize criterions (losses)"""

    @abstractmethod
    def init_optimizers(self):
        """Initialize optimizers"""

    def init_metrics(self):
        self.training_metrics = TrainingMetricsLite(self.conf)

    def init_schedulers(self):
        self.schedulers = [get_scheduler(optim, self.conf) for optim in self.optimizers.values()]

    def _specify_device(self):
        if torch.distributed.is_initialized():
            return torch.device(f"cuda:{communication.get_local_rank()}")
        if self.conf[self.conf.mode].cuda:
            return torch.device('cuda:0')
        raise RuntimeError("ganslate_b200 has no CPU path: set `cuda: True` (the hot path is sm_100a CUDA only)")

    @abstractmethod
    def set_input(self, input):
        """Unpack input data from the dataloader."""

    @abstractmethod
    def forward(self):
        """Run forward pass."""

    @abstractmethod
    def optimize_parameters(self):
        """Calculate losses, gradients, and update network weights; called in every training iteration"""

    def setup(self):
        """base.py:108-153: networks, criterions, optimizers, metrics, schedulers, checkpoint, DDP."""
        if self.conf[self.conf.mode].mixed_precision:
            raise NotImplementedError("Apex AMP is not used on the B200 path: the kernels already compute in "
                                      "bf16 with fp32 accumulation. Set mixed_precision: False.")
        self.init_networks()
        if self.is_train:
            self.init_criterions()
            self.init_optimizers()
            self.init_metrics()
            self.init_schedulers()
        else:
            self.eval()
            if len(self.networks.keys()) != 1:
                raise ValueError("When inferring there should be only one network initialized - generator.")
        if self.conf[self.conf.mode].checkpointing.load_iter:
            self.load_networks(self.conf[self.conf.mode].checkpointing.load_iter)
        num_devices = int(os.environ.get('WORLD_SIZE', torch.cuda.device_count()))
        if num_devices > 1:
            self.parallelize_networks()

    # ------------------------------------------------------------------ CUDA-graph plumbing
    def make_adam(self, params, lr, betas):
        """Adam as in the reference (cyclegan.py:81-82): a torch.optim.Adam subclass (same state_dict layout, same
        scheduler interface) whose step() is one multi-tensor sm_100a launch per 96 parameters.  With CUDA graphs the
        learning rate lives in a device scalar the captured launch reads.  `train.fused_adam: False` selects
        torch's own implementation."""
        if not bool(self.conf[self.conf.mode].get("fused_adam", True)):
            if self.use_cuda_graph:
                return torch.optim.Adam(params, lr=torch.tensor(float(lr), device=self.device), betas=betas,
                                        capturable=True)
            return torch.optim.Adam(params, lr=lr, betas=betas)
        from ganslate_b200.optim import FusedAdam
        if self.use_cuda_graph:
            return FusedAdam(params, lr=torch.tensor(float(lr), device=self.device), betas=betas)
        return FusedAdam(params, lr=lr, betas=betas)

    def stage_input(self, name, tensor):
        """Device-resident input. In graph mode the data is copied into a static buffer the graphs read."""
        if not self.use_cuda_graph:
            return tensor.to(self.device, non_blocking=True)
        buf = self._static.get(name)
        if buf is not None and buf.shape != tensor.shape and self._graphs:
            # a batch of another shape after capture (last partial batch, other patch size): this iteration runs
            # eagerly on the tensor itself, the captured graphs and their static buffers stay as they are
            self._eager_once = True
            return tensor.to(self.device, non_blocking=True)
        if buf is None or buf.shape != tensor.shape:
            buf = torch.empty(tensor.shape, dtype=torch.float32, device=self.device)
            self._static[name] = buf
        if self.input_copy_stream is not None and tensor.device.type == "cpu" and tensor.is_pinned():
            # opt-in (train.input_prefetch): the host -> device copy runs on its own stream into one of two staging
            # buffers, so it overlaps with whatever the compute stream still has queued (the previous iteration, when
            # the caller does not synchronise in between); the compute stream then only does a device-to-device copy
            # into the buffer the graphs read.  Two staging buffers: the copy of iteration i+1 never overwrites data
            # the compute stream has not consumed yet (its D2D copy of iteration i was enqueued before).
            stg = self._staging.setdefault(name, {"buf": [None, None], "read": [None, None], "k": 0})
            k = stg["k"] = stg["k"] ^ 1
            if stg["buf"][k] is None or stg["buf"][k].shape != tensor.shape:
                stg["buf"][k] = torch.empty(tensor.shape, dtype=torch.float32, device=self.device)
            cur = torch.cuda.current_stream()
            with torch.cuda.stream(self.input_copy_stream):
                if stg["read"][k] is not None:  # the compute stream's D2D copy out of this buffer, two iterations ago
                    self.input_copy_stream.wait_event(stg["read"][k])
                stg["buf"][k].copy_(tensor, non_blocking=True)
            cur.wait_stream(self.input_copy_stream)
            buf.copy_(stg["buf"][k], non_blocking=True)
            stg["read"][k] = torch.cuda.Event()
            stg["read"][k].record(cur)
            return buf
        buf.copy_(tensor, non_blocking=True)
        return buf

    def graph_mode(self, key):
        """True once the eager warm-up iterations are done (DDP needs 11 before capture)."""
        if not self.use_cuda_graph:
            return False
        if key == 'step':
            self._graph_calls += 1
            if self.__dict__.pop("_eager_once", False):
                # stage_input saw a shape the graphs were not captured for.  The eager iteration rebinds the entries
                # of visuals / losses / metrics to fresh tensors; the graphs keep writing the ones bound at capture:
                # remember those and bind them again before the next replay.
                if self._graphs and "_graph_bound" not in self.__dict__:
                    self._graph_bound = tuple(dict(d) for d in (self.visuals, self.losses, self.metrics))
                return False
            bound = self.__dict__.pop("_graph_bound", None)
            if bound is not None:
                for d, saved in zip((self.visuals, self.losses, self.metrics), bound):
                    d.update({k: v for k, v in saved.items() if not k.startswith("real")})
        return self._graph_calls > self.graph_warmup_iters

    def eager_stream(self):
        """Eager iterations that precede a capture run on a side stream (PyTorch's whole-network-capture recipe):
        autograd's AccumulateGrad nodes must not be bound to the legacy default stream."""
        import contextlib
        if not self.use_cuda_graph:
            return contextlib.nullcontext()
        if getattr(self, "_warm_stream", None) is None:
            self._warm_stream = torch.cuda.Stream()

        @contextlib.contextmanager
        def ctx():
            cur = torch.cuda.current_stream()
            self._warm_stream.wait_stream(cur)
            with torch.cuda.stream(self._warm_stream):
                yield
            cur.wait_stream(self._warm_stream)

        return ctx()

    def run_graphed(self, name, fn):
        g = self._graphs.get(name)
        if g is None:
            from ganslate_b200 import _cabi
            if not self._graphs:
                # drop the last eager iteration's autograd graph (kept alive by losses / visuals)
                for d in (self.losses, self.metrics):
                    for k in list(d):
                        d[k] = None
                for k in list(self.visuals):
                    if not k.startswith('real'):
                        self.visuals[k] = None
                self.pred_real = self.pred_fake = None
                import gc
                gc.collect()
            n0 = _cabi.lib().gb_launch_count()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            pool = next(iter(self._graphs.values())).pool() if self._graphs else None
            with torch.cuda.graph(g, pool=pool):
                fn()
            self._graphs[name] = g
            self.graph_launches_per_step += _cabi.lib().gb_launch_count() - n0
        g.replay()

    # ---- two-stream execution (train.multi_stream, opt-in; GB_MULTI_STREAM=1 sets the default) -----------------
    # The A -> B -> A and B -> A -> B cycles are independent until the losses are summed, and so are the two
    # discriminators.  With `train.multi_stream` each chain is ENQUEUED on its own CUDA stream (forked from and
    # joined to the current stream, so the CUDA-graph capture records two parallel branches): a kernel of one chain
    # fills the SMs the other chain's kernel leaves idle (second wave of a 256-tile convolution, latency-bound
    # InstanceNorm launches, everything at batch 1).  Autograd replays every node's backward on the stream its
    # forward ran on, so the backward pass forks the same way.  Every network's packed weights are refreshed on the
    # parent stream first (a chain must not launch the pack kernel the other chain depends on).
    def _streams(self, tag="chain"):
        """The two child streams of the CURRENT stream for `tag` (None: single-stream execution).  Keyed by the parent, so
        forks nest: a discriminator's real / fake passes fork again inside the stream of that discriminator."""
        default = os.environ.get("GB_MULTI_STREAM", "0") == "1"
        if not bool(self.conf.train.get("multi_stream", default)) or self.device.type != "cuda":
            return None
        from ganslate_b200 import ops
        if ops.DIRECT_PARAM_GRAD:
            return None   # (that opt-in mode adds into param.grad outside autograd's stream bookkeeping: single stream)
        pool = self.__dict__.setdefault("_chain_streams", {})
        key = (torch.cuda.current_stream().cuda_stream, tag)
        st = pool.get(key)
        if st is None:
            st = pool[key] = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        return st

    def _prepack(self, names):
        from ganslate_b200 import ops
        for n in names:
            net = self.networks[n]
            ops.ensure_packed(net.module if isinstance(net, DistributedDataParallel) else net)

    def _fork_join(self, fn1, fn2, tag="chain"):
        """Run fn1 and fn2 on two child streams of the current stream (back to back without `train.multi_stream`)."""
        st = self._streams(tag)
        if st is None:
            return fn1(), fn2()
        cur = torch.cuda.current_stream()
        st[0].wait_stream(cur)
        st[1].wait_stream(cur)
        with torch.cuda.stream(st[0]):
            r1 = fn1()
        with torch.cuda.stream(st[1]):
            r2 = fn2()
        cur.wait_stream(st[0])
        cur.wait_stream(st[1])
        for r in (r1, r2):
            for t in (r if isinstance(r, (tuple, list)) else (r,)):
                if torch.is_tensor(t):
                    t.record_stream(cur)
        return r1, r2

    def backward(self, loss, optimizer=None, retain_graph=False, loss_id=0):
        loss.backward(retain_graph=retain_graph)

    def parallelize_networks(self):
        """base.py:172-189: one DistributedDataParallel wrapper per network, broadcast_buffers=False; the
        bucketed NCCL all-reduce of the gradients overlaps with the rest of backward."""
        if torch.distributed.is_initialized() and self.use_cuda_graph and not self.graph_sync:
            self.logger.warning("%s does not synchronise gradients between CUDA-graph segments: train.cuda_graph is "
                                "switched off for the data-parallel run (eager DistributedDataParallel)", type(self).__name__)
            self.use_cuda_graph = False
        if torch.distributed.is_initialized() and self.use_cuda_graph:
            # graph-replayed steps: explicit flat-bucket all-reduce per optimizer group instead of DDP's hooks
            # (utils/grad_sync.py); same semantics -- rank-0 parameters at start, gradients averaged over ranks
            from ganslate_b200.utils.grad_sync import FlatGradSync
            self.grad_syncs = {}
            for opt_name, optim in self.optimizers.items():
                params = [p for g in optim.param_groups for p in g['params']]
                self.grad_syncs[opt_name] = FlatGradSync(params, self.device)
                self.grad_syncs[opt_name].broadcast_parameters()
            return
        for name in self.networks.keys():
            if torch.distributed.is_initialized():
                from ganslate_b200 import ops
                if ops.DIRECT_PARAM_GRAD:
                    raise RuntimeError("GB_DIRECT_PARAM_GRAD=1 bypasses autograd's gradient hooks, which DistributedDataParallel "
                                       "relies on: use it with train.cuda_graph (flat-bucket all-reduce) or on one GPU")
                self.networks[name] = DistributedDataParallel(self.networks[name], device_ids=[self.device],
                                                              output_device=self.device, broadcast_buffers=False)
            elif self.conf[self.conf.mode].cuda and torch.cuda.device_count() > 1 and "WORLD_SIZE" in os.environ:
                raise RuntimeError("Multi-GPU runs must be launched in distributed mode (torchrun).")

    def update_learning_rate(self):
        for scheduler in self.schedulers:
            scheduler.step()
        if self.use_cuda_graph:
            # keep lr a device tensor (graphs read it by address); LambdaLR assigns python floats
            for optim in self.optimizers.values():
                for group in optim.param_groups:
                    if not torch.is_tensor(group['lr']):
                        t = group.setdefault('_lr_tensor', torch.tensor(float(group['lr']), device=self.device))
                        t.fill_(float(group['lr']))
                        group['lr'] = t

    def save_checkpoint(self, iter_idx):
        """base.py:226-251 -- same file layout ({name: state_dict}, optimizer_G, optimizer_D)."""
        checkpoint = {}
        path = Path(self.output_dir) / f"checkpoints/{iter_idx}.pth"
        path.parent.mkdir(parents=True, exist_ok=True)
        for name, net in self.networks.items():
            checkpoint[name] = (net.module if isinstance(net, DistributedDataParallel) else net).state_dict()
        checkpoint['optimizer_G'] = self.optimizers['G'].state_dict()
        checkpoint['optimizer_D'] = self.optimizers['D'].state_dict()
        torch.save(checkpoint, path)

    def load_networks(self, iter_idx):
        path = Path(self.output_dir).resolve() / f"checkpoints/{iter_idx}.pth"
        checkpoint = torch.load(path, map_location=self.device)
        for name in self.networks.keys():
            net = self.networks[name]
            (net.module if isinstance(net, DistributedDataParallel) else net).load_state_dict(checkpoint[name])
        if self.is_train and self.conf[self.conf.mode].checkpointing.load_optimizers:
            self.optimizers['G'].load_state_dict(checkpoint['optimizer_G'])
            self.optimizers['D'].load_state_dict(checkpoint['optimizer_D'])

    def set_requires_grad(self, networks, requires_grad=False):
        if not isinstance(networks, list):
            networks = [networks]
        for net in networks:
            if net is not None:
                for param in net.parameters():
                    param.requires_grad = requires_grad

    def eval(self):
        for name in self.networks.keys():
            self.networks[name].eval()

    def infer(self, input):
        generator = 'G' if 'G' in self.networks.keys() else 'G_AB'
        with torch.no_grad():
            return self.networks[generator].forward(input)

    def get_loggable_data(self):
        learning_rates = {f"lr_{name}": optim.param_groups[0]['lr'] for name, optim in self.optimizers.items()}
        return learning_rates, self.losses, self.visuals, self.metrics
