"""Flat-bucket gradient all-reduce for the CUDA-graph training path.

The reference wraps every network in DistributedDataParallel (ganslate/nn/gans/base.py:172-189) and that wrap is
kept for eager execution.  DDP's reducer hooks do not survive being captured into the replayed graphs on this
stack (the capture dead-locks), so when `train.cuda_graph` is on the same average-all-reduce is issued explicitly:
after a backward graph has run, the gradients of one optimizer group are packed into one fp32 buffer and reduced
with NCCL on a side stream, overlapping with the next graph (the discriminator phase); the optimizer step waits for
it.  Semantics are DDP's: mean over ranks, parameters broadcast from rank 0 at start.
"""
import os

import torch
import torch.distributed as dist

# Opt-in (GB_SYNC_BF16=1): the bucket travels as bf16 -- half the bytes on the wire, which matters once the bucket is
# hundreds of MB (Pix2Pix U-Net: 669 MB of fp32 generator gradients per step) and little of it can be hidden.  Each
# rank's contribution is pre-divided by the world size and rounded to bf16 once; NCCL sums in bf16.  The averaged
# gradient then carries ~3 significant digits, as with DDP's bf16 compression hook.
SYNC_BF16 = os.environ.get("GB_SYNC_BF16", "0") == "1"


class FlatGradSync:

    def __init__(self, params, device, dtype=None):
        seen, self.params = set(), []
        for p in params:
            if id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        self.device = device
        self.dtype = dtype if dtype is not None else (torch.bfloat16 if SYNC_BF16 else torch.float32)
        self.flat = torch.empty(sum(p.numel() for p in self.params), dtype=self.dtype, device=device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        # NCCL runs on a side stream; with CPU tensors (gloo: the host-logic tests) everything is synchronous
        self.stream = torch.cuda.Stream(device=device) if torch.device(device).type == "cuda" else None
        self.world = dist.get_world_size()
        self._pending = False

    def broadcast_parameters(self):
        for p in self.params:
            dist.broadcast(p.data, src=0)
        # the write went through `.data`: bump the version counters, ConvOp's packed-weight cache keys on them (a
        # network evaluated before this call -- CUT's channel probe -- would otherwise keep its pre-broadcast copies)
        torch.autograd.graph.increment_version(self.params)

    def launch(self):
        """Call after backward on the compute stream: pack and start the all-reduce on the side stream."""
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, grads)
        # fp32 bucket: average after the sum.  bf16 bucket: scale BEFORE the sum so that it stays in range (exact for
        # the power-of-two world sizes of one node: a change of exponent)
        pre = self.dtype != torch.float32

        def reduce():
            if pre:
                self.flat.mul_(1.0 / self.world)
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if not pre:
                self.flat.mul_(1.0 / self.world)

        if self.stream is None:
            reduce()
        else:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                reduce()
        self._pending = True

    def finish(self):
        """Call before the optimizer step: wait for the reduction and write the averaged gradients back."""
        if not self._pending:
            return
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        grads = [p.grad for p in self.params if p.grad is not None]
        views = [v for p, v in zip(self.params, self.views) if p.grad is not None]
        torch._foreach_copy_(grads, views)
        self._pending = False
