"""TEST INFRASTRUCTURE ONLY -- CPU fp32 oracle of the 3-D part of ganslate's hot path: Vnet3D (partially
invertible V-Net), PatchGAN3D, and one RevGAN / 3-D CycleGAN training iteration.

Same rules as oracle/torch_oracle.py: plain PyTorch on the CPU, imported only by tests/, smoke() and bench.py's CPU
legs.  Module attribute names follow the reference so that `state_dict()` keys are identical and weights can be
exchanged with both the reference modules (tests/test_oracle.py pins this file to them when /root/reference is
present) and ganslate_b200's modules.

memcnn (third-party, unpinned, absent): AdditiveCoupling is restated from its published algorithm,
y1 = x1 + Fm(x2), y2 = x2 + Gm(y1); inverse x2 = y2 - Gm(y1), x1 = y1 - Fm(x2) (call sites
ganslate/nn/invertible.py:15-19).  No reference test holds golden values for it: parity unpinned at that boundary;
the pin is invertibility G^-1(G(x)) = x and equality with the reference modules run over the same stand-in.
"""
import copy
import itertools
from types import SimpleNamespace

import torch
from torch import nn

from oracle.torch_oracle import OracleImagePool, adversarial_lsgan, cyclegan_losses, init_weights


def _in3d(c):
    return nn.InstanceNorm3d(c)  # ganslate/nn/utils.py:62-68: eps 1e-5, affine=False, no running stats


class _SepConv3d(nn.Module):
    """ganslate/nn/separable.py:5-40: (1, k, k) in-plane convolution, then (k, 1, 1) through-plane convolution."""

    def __init__(self, cin, cout, k, stride=1, padding=0, bias=True, transposed=False):
        super().__init__()
        conv = nn.ConvTranspose3d if transposed else nn.Conv3d
        names = ("conv_transp_depthwise", "conv_transp_pointwise") if transposed else ("conv_depthwise", "conv_pointwise")
        self._names = names
        setattr(self, names[0], conv(cin, cout, (1, k, k), stride=(1, stride, stride), padding=(0, padding, padding),
                                     bias=bias))
        setattr(self, names[1], conv(cout, cout, (k, 1, 1), stride=(stride, 1, 1), padding=(padding, 0, 0), bias=bias))

    def forward(self, x):
        return getattr(self, self._names[1])(getattr(self, self._names[0])(x))


def _conv3d(sep, cin, cout, k, transposed=False, **kw):
    """ganslate/nn/utils.py:39-50: nn.Conv3d / nn.ConvTranspose3d or their separable counterparts."""
    if sep:
        return _SepConv3d(cin, cout, k, transposed=transposed, **kw)
    return (nn.ConvTranspose3d if transposed else nn.Conv3d)(cin, cout, k, **kw)


def _cnp(cin, cout, k, sep=False, **kw):
    """[conv, norm, PReLU] group used all over vnet3d.py (e.g. :186-189, :262-267)."""
    return nn.Sequential(_conv3d(sep, cin, cout, k, bias=True, **kw), _in3d(cout), nn.PReLU(cout))


class _Coupling(nn.Module):
    """memcnn.AdditiveCoupling stand-in (module names Fm / Gm as in memcnn)."""

    def __init__(self, fm):
        super().__init__()
        self.Fm, self.Gm = fm, copy.deepcopy(fm)

    def forward(self, x):
        x1, x2 = torch.chunk(x, 2, dim=1)
        y1 = x1 + self.Fm(x2)
        return torch.cat([y1, x2 + self.Gm(y1)], dim=1)

    def inverse(self, y):
        y1, y2 = torch.chunk(y, 2, dim=1)
        x2 = y2 - self.Gm(y1)
        return torch.cat([y1 - self.Fm(x2), x2], dim=1)


class _Wrapper(nn.Module):
    """memcnn.InvertibleModuleWrapper stand-in: same values and gradients, activations simply kept."""

    def __init__(self, fn):
        super().__init__()
        self._fn = fn


class _InvBlock(nn.Module):  # ganslate/nn/invertible.py:8-24
    def __init__(self, block):
        super().__init__()
        self.invertible_block = _Wrapper(_Coupling(copy.deepcopy(block)))


class _InvSequence(nn.Module):  # ganslate/nn/invertible.py:27-48
    def __init__(self, block, n):
        super().__init__()
        self.sequence = nn.Sequential(*[_InvBlock(block) for _ in range(n)])

    def forward(self, x, inverse=False):
        blocks = reversed(self.sequence) if inverse else self.sequence
        for b in blocks:
            fn = b.invertible_block._fn
            x = fn.inverse(x) if inverse else fn(x)
        return x


class _In(nn.Module):  # vnet3d.py:151-167
    def __init__(self, cin, cout, sep=False):
        super().__init__()
        self.n_repeats = cout // cin
        self.conv1, self.bn1, self.relu = _conv3d(sep, cin, cout, 5, padding=2, bias=True), _in3d(cout), nn.PReLU(cout)

    def forward(self, x):
        return self.relu(self.bn1(self.conv1(x)) + x.repeat(1, self.n_repeats, 1, 1, 1))


class _Down(nn.Module):  # vnet3d.py:170-202
    def __init__(self, cin, n_blocks, use_inverse, sep=False):
        super().__init__()
        c = 2 * cin
        self.down_conv_ab = _cnp(cin, c, 2, sep, stride=2)
        if use_inverse:
            self.down_conv_ba = _cnp(cin, c, 2, sep, stride=2)
        self.core = _InvSequence(_cnp(c // 2, c // 2, 5, sep, padding=2), n_blocks)
        self.relu = nn.PReLU(c)

    def forward(self, x, inverse=False):
        down = (self.down_conv_ba if inverse else self.down_conv_ab)(x)
        return self.relu(self.core(down, inverse) + down)


class _Up(nn.Module):  # vnet3d.py:205-240
    def __init__(self, cin, cout, n_blocks, use_inverse, sep=False):
        super().__init__()

        def up():
            return nn.Sequential(_conv3d(sep, cin, cout // 2, 2, transposed=True, stride=2, bias=True), _in3d(cout // 2),
                                 nn.PReLU(cout // 2))

        self.up_conv_ab = up()
        if use_inverse:
            self.up_conv_ba = up()
        self.core = _InvSequence(_cnp(cout // 2, cout // 2, 5, sep, padding=2), n_blocks)
        self.relu = nn.PReLU(cout)

    def forward(self, x, skip, inverse=False):
        xcat = torch.cat(((self.up_conv_ba if inverse else self.up_conv_ab)(x), skip), 1)
        return self.relu(self.core(xcat, inverse) + xcat)


class _Out(nn.Module):  # vnet3d.py:243-259
    def __init__(self, cin, cout, sep=False):
        super().__init__()
        self.conv1, self.bn1, self.relu1 = _conv3d(sep, cin, cin, 5, padding=2, bias=True), _in3d(cin), nn.PReLU(cin)
        self.conv2, self.tanh = _conv3d(sep, cin, cout, 1), nn.Tanh()

    def forward(self, x):
        return self.tanh(self.conv2(self.relu1(self.bn1(self.conv1(x)))))


class OracleVnet3D(nn.Module):
    """ganslate/nn/generators/vnet/vnet3d.py:27-148 (norm_type 'instance')."""

    def __init__(self, in_channels, out_channels, first_layer_channels=16, down_blocks=(1, 2, 3, 2),
                 up_blocks=(2, 2, 1, 1), use_inverse=True, is_separable=False):
        super().__init__()
        f = first_layer_channels
        sep = is_separable
        self.use_inverse = use_inverse
        self.in_ab = _In(in_channels, f, sep)
        if use_inverse:
            self.in_ba = _In(in_channels, f, sep)
        self.out_ab = _Out(2 * f, out_channels, sep)
        if use_inverse:
            self.out_ba = _Out(2 * f, out_channels, sep)
        self.downs = nn.ModuleList([_Down(f * 2**i, n, use_inverse, sep) for i, n in enumerate(down_blocks)])
        self.encoder = nn.ModuleList([self.in_ab]).extend(self.downs)  # :88
        factors = [2 * 2**i for i in reversed(range(len(down_blocks)))]   # :91
        ups = [_Up(f * factors[0], f * factors[0], up_blocks[0], use_inverse, sep)]
        for i, n in enumerate(up_blocks[1:]):
            ups.append(_Up(f * factors[i], f * factors[i + 1], n, use_inverse, sep))
        self.ups = nn.ModuleList(ups)

    def forward(self, x, inverse=False):  # :107-148
        if inverse and not self.use_inverse:
            raise ValueError("inverse pass requested but use_inverse is off")
        out1 = (self.in_ba if inverse else self.in_ab)(x)
        downs = []
        for d in self.downs:
            downs.append(d(downs[-1] if downs else out1, inverse))
        rev = downs[::-1]
        out = rev[0]
        for i, u in enumerate(self.ups):
            out = u(out, out1 if i == len(self.ups) - 1 else rev[i + 1], inverse)
        return (self.out_ba if inverse else self.out_ab)(out)


class _ResBlock3d(nn.Module):  # ganslate/nn/generators/resnet/resnet3d.py:72-91
    def __init__(self, c):
        super().__init__()
        self.conv_block = nn.Sequential(nn.ReplicationPad3d(1), nn.Conv3d(c, c, 3, bias=True), _in3d(c), nn.ReLU(True),
                                        nn.ReplicationPad3d(1), nn.Conv3d(c, c, 3, bias=True), _in3d(c))

    def forward(self, x):
        return x + self.conv_block(x)


class OracleResnet3D(nn.Module):
    """ganslate/nn/generators/resnet/resnet3d.py:14-69 (norm_type 'instance')."""

    def __init__(self, in_channels, out_channels, n_residual_blocks=9):
        super().__init__()
        L = [nn.ReplicationPad3d(3), nn.Conv3d(in_channels, 64, 7, bias=True), _in3d(64), nn.ReLU(True)]
        c = 64
        for _ in range(2):  # :32-41
            L += [nn.Conv3d(c, 2 * c, 3, stride=2, padding=1, bias=True), _in3d(2 * c), nn.ReLU(True)]
            c *= 2
        L += [_ResBlock3d(c) for _ in range(n_residual_blocks)]  # :44-45
        for _ in range(2):  # :48-61 (ConvTranspose3d keeps torch's default bias=True)
            L += [nn.ConvTranspose3d(c, c // 2, 3, stride=2, padding=1, output_padding=1), _in3d(c // 2), nn.ReLU(True)]
            c //= 2
        L += [nn.ReplicationPad3d(3), nn.Conv3d(64, out_channels, 7, bias=True), nn.Tanh()]  # :64
        self.model = nn.Sequential(*L)

    def forward(self, x):
        return self.model(x)


class OraclePiresnet3D(nn.Module):
    """ganslate/nn/generators/resnet/piresnet3d.py:28-119 (norm_type 'instance')."""

    def __init__(self, in_channels, out_channels, depth, first_layer_channels=64, use_inverse=True):
        super().__init__()
        c = first_layer_channels
        self.use_inverse = use_inverse

        def down():  # :58-75
            return nn.Sequential(nn.ReplicationPad3d(2), nn.Conv3d(in_channels, c, 5, bias=True), _in3d(c), nn.ReLU(True),
                                 nn.Conv3d(c, 2 * c, 3, stride=2, padding=1, bias=True), _in3d(2 * c), nn.ReLU(True))

        def up():  # :77-87 (the last convolution keeps torch's default bias=True)
            return nn.Sequential(nn.ConvTranspose3d(2 * c, c, 3, stride=2, padding=1, output_padding=1, bias=True),
                                 _in3d(c), nn.ReLU(True), nn.ReplicationPad3d(2), nn.Conv3d(c, out_channels, 5), nn.Tanh())

        self.downconv_ab, self.upconv_ab = down(), up()
        if use_inverse:
            self.downconv_ba, self.upconv_ba = down(), up()
        inv = nn.Sequential(_in3d(c), nn.ReplicationPad3d(1), nn.Conv3d(c, c, 3, bias=True), _in3d(c), nn.ReLU(True))  # :114-119
        self.core = _InvSequence(inv, depth)

    def forward(self, x, inverse=False):  # :89-111
        if inverse and not self.use_inverse:
            raise ValueError("inverse pass requested but use_inverse is off")
        down, up = (self.downconv_ba, self.upconv_ba) if inverse else (self.downconv_ab, self.upconv_ab)
        return up(self.core(down(x), inverse))


class OraclePatchGAN3D(nn.Module):
    """ganslate/nn/discriminators/patchgan/patchgan3d.py:17-65"""

    def __init__(self, in_channels, ndf=64, n_layers=3, kernel_size=(4, 4, 4)):
        super().__init__()
        k = tuple(kernel_size)
        L = [nn.Conv3d(in_channels, ndf, k, stride=2, padding=1), nn.LeakyReLU(0.2, True)]
        mult = 1
        for n in range(1, n_layers):
            prev, mult = mult, min(2**n, 8)
            L += [nn.Conv3d(ndf * prev, ndf * mult, k, stride=2, padding=1), _in3d(ndf * mult), nn.LeakyReLU(0.2, True)]
        prev, mult = mult, min(2**n_layers, 8)
        L += [nn.Conv3d(ndf * prev, ndf * mult, k, stride=1, padding=1), _in3d(ndf * mult), nn.LeakyReLU(0.2, True)]
        L += [nn.Conv3d(ndf * mult, 1, k, stride=1, padding=1)]
        self.model = nn.Sequential(*L)

    def forward(self, x):
        return self.model(x)


def default_3d_conf(**kw):
    c = dict(lambda_AB=10.0, lambda_BA=10.0, lambda_identity=0.0, lr_G=2e-4, lr_D=2e-4, beta1=0.5, beta2=0.999,
             pool_size=50, in_channels=1, out_channels=1, first_layer_channels=16, down_blocks=(1, 2, 3, 2),
             up_blocks=(2, 2, 1, 1), ndf=64, n_layers=3, kernel_size=(4, 4, 4))
    c.update(kw)
    return SimpleNamespace(**c)


class OracleRevGAN:
    """One iteration of ganslate/nn/gans/unpaired/revgan.py:89-212: ONE partially invertible generator used in
    both directions (`inverse=True` = B->A), discriminator inputs swapped in backward_G exactly as the reference
    does (:196-197: pred_B = D_B(fake_A), pred_A = D_A(fake_B))."""

    def __init__(self, conf=None, seed=0):
        self.conf = c = conf or default_3d_conf()
        torch.manual_seed(seed)
        self.networks = {}
        for name in ("G", "D_B", "D_A"):  # revgan.py:50, base.py:51-67
            if name == "G" and getattr(c, "piresnet_depth", 0):  # brats revgan.yaml:27-33
                net = OraclePiresnet3D(c.in_channels, c.out_channels, c.piresnet_depth, c.first_layer_channels, True)
            elif name == "G":
                net = OracleVnet3D(c.in_channels, c.out_channels, c.first_layer_channels, c.down_blocks, c.up_blocks,
                                   use_inverse=True)
            else:
                net = OraclePatchGAN3D(c.in_channels, c.ndf, c.n_layers, c.kernel_size)
            self.networks[name] = init_weights(net)
        n = self.networks
        self.optimizers = {  # revgan.py:67-79
            "G": torch.optim.Adam(n["G"].parameters(), lr=c.lr_G, betas=(c.beta1, c.beta2)),
            "D": torch.optim.Adam(itertools.chain(n["D_B"].parameters(), n["D_A"].parameters()), lr=c.lr_D,
                                  betas=(c.beta1, c.beta2)),
        }
        self.fake_A_pool, self.fake_B_pool = OracleImagePool(c.pool_size), OracleImagePool(c.pool_size)
        self.visuals, self.losses = {}, {}

    def forward(self):  # revgan.py:122-150
        G, v = self.networks["G"], self.visuals
        v["fake_B"] = G(v["real_A"])
        v["rec_A"] = G(v["fake_B"], inverse=True)
        v["fake_A"] = G(v["real_B"], inverse=True)
        v["rec_B"] = G(v["fake_A"])
        v["idt_A"] = v["idt_B"] = None
        if self.conf.lambda_identity > 0:
            v["idt_B"] = G(v["real_B"])
            v["idt_A"] = G(v["real_A"], inverse=True)

    def backward_G(self):  # revgan.py:188-212
        n, v, c = self.networks, self.visuals, self.conf
        self.losses["G_AB"] = adversarial_lsgan(n["D_B"](v["fake_A"]), True)
        self.losses["G_BA"] = adversarial_lsgan(n["D_A"](v["fake_B"]), True)
        lg = cyclegan_losses(v, c.lambda_AB, c.lambda_BA, c.lambda_identity)
        self.losses.update(lg)
        (sum(lg.values()) + self.losses["G_AB"] + self.losses["G_BA"]).backward()

    def backward_D(self, name):  # revgan.py:152-186
        v = self.visuals
        if name == "D_B":
            real, fake = v["real_B"], self.fake_B_pool.query(v["fake_B"])
        else:
            real, fake = v["real_A"], self.fake_A_pool.query(v["fake_A"])
        pr, pf = self.networks[name](real), self.networks[name](fake.detach())
        self.losses[name] = adversarial_lsgan(pr, True) + adversarial_lsgan(pf, False)
        self.losses[name].backward(retain_graph=True)

    def optimize_parameters(self, real_A, real_B, step_optimizers=True):  # revgan.py:89-120
        self.visuals["real_A"], self.visuals["real_B"] = real_A, real_B
        ds = [self.networks["D_B"], self.networks["D_A"]]
        self.forward()
        for d in ds:
            d.requires_grad_(False)
        self.optimizers["G"].zero_grad(set_to_none=True)
        self.backward_G()
        if step_optimizers:
            self.optimizers["G"].step()
        for d in ds:
            d.requires_grad_(True)
        self.optimizers["D"].zero_grad(set_to_none=True)
        self.backward_D("D_B")
        self.backward_D("D_A")
        if step_optimizers:
            self.optimizers["D"].step()
        return {k: float(v.detach()) for k, v in self.losses.items() if v is not None}


def synthetic_volume(batch, channels, depth, size, seed=1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = torch.rand((batch, channels, depth, size, size), generator=g) * 2 - 1
    b = torch.rand((batch, channels, depth, size, size), generator=g) * 2 - 1
    return a, b
