"""Additive-coupling containers with the module / state_dict layout of ganslate/nn/invertible.py:8-48 on top of
memcnn (`...core.sequence.{i}.invertible_block._fn.{Fm,Gm}.{0,2}.*`).

memcnn.AdditiveCoupling: y1 = x1 + Fm(x2); y2 = x2 + Gm(y1) on a channel split in halves, inverse
x2 = y2 - Gm(y1); x1 = y1 - Fm(x2) (memcnn is an unpinned dependency absent from this image; its published
algorithm is restated).  memcnn.InvertibleModuleWrapper only changes WHEN activations are
stored (it frees the input and recomputes it in backward); on a 180 GB part the activations are simply kept, so
`keep_input` is accepted and ignored -- values and gradients are identical.

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

    def gb_run(self, tape, x, inverse=False):
        blocks = list(reversed(self.sequence)) if inverse else list(self.sequence)
        for blk in blocks:
            fn = blk.invertible_block._fn
            x = coupling_inverse(tape, x, fn) if inverse else coupling_forward(tape, x, fn)
        return x


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
