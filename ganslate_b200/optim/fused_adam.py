"""Adam with the whole parameter list updated by one or two sm_100a launches (csrc/adam.cu).

A subclass of torch.optim.Adam: construction, param_groups, `state_dict()` / `load_state_dict()` (so the
checkpoint layout of ganslate/nn/gans/base.py:226-251 is unchanged) and the LambdaLR scheduler of
ganslate/nn/utils.py:83-99 work as before; only `step()` is replaced.  Per parameter the state holds `exp_avg`,
`exp_avg_sq` and `step` exactly like torch's implementation (`step` is one device scalar per group shared by all of
its parameters, which is what makes the launch CUDA-graph replayable)."""
import ctypes as C

import torch

from .. import _cabi


class FusedAdam(torch.optim.Adam):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False)
        self._lr_dev = {}

    def _group_lr(self, gi, group, device):
        """Device scalar holding the group's learning rate (a tensor lr is used as is)."""
        lr = group["lr"]
        if torch.is_tensor(lr) and lr.is_cuda:
            return lr if lr.dtype == torch.float32 else lr.float()
        ent = self._lr_dev.get(gi)
        val = float(lr)
        if ent is None or ent[0].device != device:
            ent = [torch.tensor(val, dtype=torch.float32, device=device), val]
            self._lr_dev[gi] = ent
        elif ent[1] != val:
            ent[0].fill_(val)
            ent[1] = val
        return ent[0]

    def state_dict(self):
        """torch.optim.Adam's layout with ONE `step` PER PARAMETER (a CPU scalar, as torch writes it): the shared device
        scalar of a group must not survive into a checkpoint -- torch.optim.Adam (the reference, or
        `train.fused_adam: False`) would add 1 to the shared tensor once per parameter and step."""
        sd = super().state_dict()
        read = {}   # one device -> host read per shared scalar, not per parameter
        for st in sd["state"].values():
            s = st.get("step")
            if torch.is_tensor(s):
                if id(s) not in read:
                    read[id(s)] = float(s)
                st["step"] = torch.tensor(read[id(s)], dtype=torch.float32)
        return sd

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _cabi.lib()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            dev = params[0].device
            if not params[0].is_cuda:
                raise RuntimeError("ganslate_b200.optim.FusedAdam: parameters must live on a CUDA device")
            step_t = None
            for p in params:
                st = self.state[p]
                if len(st) == 0:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if step_t is None:
                    s = st.get("step")
                    if torch.is_tensor(s) and s.is_cuda and s.dtype == torch.float32:
                        step_t = s
            if step_t is None:
                s0 = self.state[params[0]].get("step", 0.0)  # e.g. a CPU scalar restored from a torch checkpoint
                step_t = torch.tensor(float(s0), dtype=torch.float32, device=dev)
            for p in params:
                self.state[p]["step"] = step_t
            step_t.add_(1.0)
            lr_t = self._group_lr(gi, group, dev)
            b1, b2 = group["betas"]
            batch = _cabi.AdamBatch()
            batch.lr, batch.step = lr_t.data_ptr(), step_t.data_ptr()
            batch.beta1, batch.beta2, batch.eps = float(b1), float(b2), float(group["eps"])
            n = 0
            for p in params:
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous() or not p.is_contiguous():
                    raise RuntimeError("FusedAdam expects contiguous fp32 parameters and gradients")
                st = self.state[p]
                it = batch.item[n]
                it.p, it.g, it.m, it.v = p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                it.n = p.numel()
                it.vec4 = 1 if all(a % 16 == 0 for a in (it.p, it.g, it.m, it.v)) else 0
                n += 1
                if n == _cabi.GB_ADAM_BATCH:
                    batch.count = n
                    _cabi.check(lib.gb_adam_multi(C.byref(batch), stream), "gb_adam_multi")
                    n = 0
            if n:
                batch.count = n
                _cabi.check(lib.gb_adam_multi(C.byref(batch), stream), "gb_adam_multi")
            # the kernel wrote the parameters through raw pointers: tell autograd (and the packed-weight cache of
            # ganslate_b200.ops, which keys on Tensor._version) that they changed
            torch.autograd.graph.increment_version(params)
        return loss
