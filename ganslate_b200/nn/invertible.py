"""Additive-coupling containers with the module / state_dict layout of ganslate/nn/invertible.py:8-48 on top of
memcnn (`...core.sequence.{i}.invertible_block._fn.{Fm,Gm}.{0,2}.*`).

memcnn.AdditiveCoupling: y1 = x1 + Fm(x2); y2 = x2 + Gm(y1) on a channel split in halves, inverse
x2 = y2 - Gm(y1); x1 = y1 - Fm(x2) (memcnn is an unpinned dependency absent from this image; its published
algorithm is restated).

memcnn.InvertibleModuleWrapper decides WHEN activations exist: it runs the coupling without recording anything,
frees the block's INPUT unless `keep_input` (ganslate/nn/invertible.py:41-46 forces keep_input on the first block of
a sequence, whose input other layers still read), and in backward rebuilds that input from the block's OUTPUT with
the inverse coupling, re-runs the coupling with recording on and back-propagates through it.  `recompute_coupling`
below does exactly that on the tape (`use_memory_saving=True` -> `keep_input=False`, vnet3d.py:52): per sequence
only the first block's input and the last block's output stay allocated between forward and backward -- every
other block input and every intermediate of every block (raw convolution outputs, statistics) is transient.
Values: the rebuilt input is `bf16(y - F(.))` of `y = bf16(x + F(.))`, i.e. equal to the original up to one bf16
rounding per block (the reference's fp32 rebuild is off by one fp32 rounding in the same place); the gradients of the
two modes agree to that level (tests/test_host_networks_cpu.py::test_vnet3d_memory_saving_*, tests/test_3d_gpu.py).

The classes hold parameters only; the compute is `coupling_forward` / `coupling_inverse` below, built from the fused
step primitives on channel-slice views (no split / cat copies)."""
from copy import deepcopy

from torch import nn

from ganslate_b200.nn import layers
from ganslate_b200._cabi import ACT_PRELU


class AdditiveCoupling(nn.Module):

    def __init__(self, Fm, Gm=None):
        super().__init__()
        self.Fm = Fm
        self.Gm = deepcopy(Fm) if Gm is None else Gm


class InvertibleModuleWrapper(nn.Module):

    def __init__(self, fn, keep_input=False, keep_input_inverse=False, disable=False):
        super().__init__()
        self._fn = fn
        self.keep_input, self.keep_input_inverse, self.disable = keep_input, keep_input_inverse, disable


class InvertibleBlock(nn.Module):

    def __init__(self, block, keep_input, disable=False):
        super().__init__()
        block = AdditiveCoupling(deepcopy(block))
        self.invertible_block = InvertibleModuleWrapper(fn=block, keep_input=keep_input, keep_input_inverse=keep_input,
                                                        disable=disable)


class InvertibleSequence(nn.Module):

    def __init__(self, block, n_blocks, keep_input, disable=False):
        super().__init__()
        self.sequence = nn.Sequential(*[InvertibleBlock(block, keep_input, disable) for _ in range(n_blocks)])

    def gb_run_coupling(self, tape, x, inverse=False):
        """(not `gb_run`: layers.run_sequence's protocol passes a border as the third argument)"""
        blocks = list(reversed(self.sequence)) if inverse else list(self.sequence)
        for i, blk in enumerate(blocks):
            wrap = blk.invertible_block
            fn = wrap._fn
            keep = wrap.keep_input_inverse if inverse else wrap.keep_input
            if tape is None or wrap.disable or keep:
                # nothing to save (no backward will run) or the caller keeps every activation
                x = coupling_inverse(tape, x, fn) if inverse else coupling_forward(tape, x, fn)
            else:
                # ganslate/nn/invertible.py:41-46: the first block of a sequence keeps its input
                x = recompute_coupling(tape, x, fn, inverse, free_input=(i > 0))
        return x


RECOMPUTE_STATS = {"blocks": 0, "rebuilt_inputs": 0}  # counters for the tests (how often the backward path below ran)


def recompute_coupling(tape, x, fn, inverse, free_input):
    """memcnn.InvertibleModuleWrapper(keep_input=False) on the tape: forward without recording; backward =
    [rebuild the input from the output with the inverse coupling] + recorded re-run + its backward."""
    fwd, inv = (coupling_inverse, coupling_forward) if inverse else (coupling_forward, coupling_inverse)
    x_st = x.st
    n_cons = x_st.consumers
    y = fwd(None, x, fn)                 # intermediates die here
    x_st.consumers = n_cons + 1
    shape = tuple(x_st.t.shape)
    if free_input:
        x_st.t = None                    # the input's memory goes back to the allocator (memcnn: storage().resize_(0))

    def bwd():
        if not y.has_grad():
            return
        g = y.st.grad
        y.st.grad = None
        RECOMPUTE_STATS["blocks"] += 1
        if x_st.t is None:
            n_y = y.st.consumers
            xr = inv(None, y, fn)        # rebuilt from the output (which the next block's backward rebuilt before)
            y.st.consumers = n_y
            assert tuple(xr.st.t.shape) == shape
            x_st.t = xr.st.t
            RECOMPUTE_STATS["rebuilt_inputs"] += 1
        sub = layers.Tape(tape.param_needs_grad, tape.input_needs_grad)
        sub.param_grads, sub.unpack, sub._queued = tape.param_grads, tape.unpack, tape._queued
        n_cons = x_st.consumers
        y2 = fwd(sub, x, fn)
        x_st.consumers = n_cons
        y2.st.grad = g
        steps, sub.steps = sub.steps, None
        while steps:
            steps.pop()()
        if "_side" in sub.__dict__:          # weight gradients on a side stream: the enclosing pass joins it
            tape.__dict__["_side"] = sub.__dict__.pop("_side")

    tape.steps.append(bwd)
    return y


def _branch(tape, src, seq, residual, out, out_scale):
    """out = residual + out_scale * act(IN(conv(head(src)))).

    V-Net (vnet3d.py:261-267): seq = [conv, norm, PReLU].  Piresnet3D (piresnet3d.py:114-119): seq = [norm,
    ReplicationPad3d, conv, norm, ReLU] -- the layers in front of the last convolution run as ordinary steps, the
    tail is fused with the coupling's add."""
    seq = list(seq)
    k = max(i for i, m in enumerate(seq) if isinstance(m, layers._ConvMixin) or hasattr(m, "gb_pair"))
    head, conv, tail = seq[:k], seq[k], seq[k + 1:]
    if len(tail) != 2 or not layers._is_norm(tail[0]):
        raise NotImplementedError("coupling branch must end with [conv, norm, activation]")
    if head:
        src = layers.run_sequence(tape, head, src)
    norm, actm = tail
    raw = layers.step_conv_any(tape, src, conv, want_stats=True)
    if isinstance(actm, layers.PReLU):
        layers.step_norm_act(tape, raw, True, ACT_PRELU, 0.0, 0, norm.eps, residual=residual, prelu=actm,
                             out_scale=out_scale, out=out)
    else:
        act_id, slope = layers._act_of(actm)
        layers.step_norm_act(tape, raw, True, act_id, slope, 0, norm.eps, residual=residual, out_scale=out_scale,
                             out=out)


def coupling_forward(tape, x, fn):
    h = x.channels // 2
    y = layers.new_like(x, x.channels)
    x1, x2, y1, y2 = x.slice(0, h), x.slice(h, h), y.slice(0, h), y.slice(h, h)
    _branch(tape, x2, fn.Fm, x1, y1, 1.0)   # y1 = x1 + Fm(x2)
    _branch(tape, y1, fn.Gm, x2, y2, 1.0)   # y2 = x2 + Gm(y1)
    return y


def coupling_inverse(tape, y, fn):
    h = y.channels // 2
    x = layers.new_like(y, y.channels)
    y1, y2, x1, x2 = y.slice(0, h), y.slice(h, h), x.slice(0, h), x.slice(h, h)
    _branch(tape, y1, fn.Gm, y2, x2, -1.0)  # x2 = y2 - Gm(y1)
    _branch(tape, x2, fn.Fm, y1, x1, -1.0)  # x1 = y1 - Fm(x2)
    return x
