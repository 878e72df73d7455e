"""TEST INFRASTRUCTURE ONLY -- CPU fp32 oracle of ganslate's training hot path.

A restatement, in plain PyTorch on the CPU, of the reference's networks, losses and one CycleGAN / Pix2Pix
training iteration.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product path (ganslate_b200/) never does.

Parity pin: the reference ships no golden vectors, so the oracle is pinned against the
reference's own modules imported in the build container (tests/test_oracle_vs_reference.py) and against the
fixtures under tests/golden/ that oracle/make_golden.py generated from those modules.

Each function cites the reference lines it restates (paths relative to the reference root).
"""
import itertools
import random
from types import SimpleNamespace

import torch
import torch.nn.functional as F
from torch import nn


# --------------------------------------------------------------------------------------------- networks
def _norm2d(c):
    # ganslate/nn/utils.py:53-59 -> nn.InstanceNorm2d defaults: eps 1e-5, affine=False, no running stats
    return nn.InstanceNorm2d(c)


class OracleResBlock(nn.Module):
    """ganslate/nn/generators/resnet/resnet2d.py:73-93"""

    def __init__(self, c):
        super().__init__()
        seq = []
        for last in (False, True):
            seq += [nn.ReflectionPad2d(1), nn.Conv2d(c, c, 3, bias=True), _norm2d(c)]
            if not last:
                seq.append(nn.ReLU(inplace=True))
        self.conv_block = nn.Sequential(*seq)

    def forward(self, x):
        return x + self.conv_block(x)


class OracleResnet2D(nn.Module):
    """ganslate/nn/generators/resnet/resnet2d.py:14-70 (norm_type 'instance' => conv bias on, :20)"""

    def __init__(self, in_channels, out_channels, n_residual_blocks=9):
        super().__init__()
        L = [nn.ReflectionPad2d(3), nn.Conv2d(in_channels, 64, 7), _norm2d(64), nn.ReLU(inplace=True)]
        c = 64
        for _ in range(2):  # :32-41
            L += [nn.Conv2d(c, 2 * c, 3, stride=2, padding=1), _norm2d(2 * c), nn.ReLU(inplace=True)]
            c *= 2
        L += [OracleResBlock(c) for _ in range(n_residual_blocks)]  # :43-44
        self.encoder = nn.ModuleList(L)  # :46 (aliases the same module objects)
        for _ in range(2):  # :48-60
            L += [nn.ConvTranspose2d(c, c // 2, 3, stride=2, padding=1, output_padding=1), _norm2d(c // 2),
                  nn.ReLU(inplace=True)]
            c //= 2
        L += [nn.ReflectionPad2d(3), nn.Conv2d(64, out_channels, 7), nn.Tanh()]  # :63
        self.model = nn.Sequential(*L)

    def forward(self, x):
        return self.model(x)


class OraclePatchGAN2D(nn.Module):
    """ganslate/nn/discriminators/patchgan/patchgan2d.py:17-66"""

    def __init__(self, in_channels, ndf=64, n_layers=3, kernel_size=(4, 4)):
        super().__init__()
        k = tuple(kernel_size)
        L = [nn.Conv2d(in_channels, ndf, k, stride=2, padding=1), nn.LeakyReLU(0.2, True)]
        mult = 1
        for n in range(1, n_layers):
            prev, mult = mult, min(2**n, 8)
            L += [nn.Conv2d(ndf * prev, ndf * mult, k, stride=2, padding=1), _norm2d(ndf * mult), nn.LeakyReLU(0.2, True)]
        prev, mult = mult, min(2**n_layers, 8)
        L += [nn.Conv2d(ndf * prev, ndf * mult, k, stride=1, padding=1), _norm2d(ndf * mult), nn.LeakyReLU(0.2, True)]
        L += [nn.Conv2d(ndf * mult, 1, k, stride=1, padding=1)]
        self.model = nn.Sequential(*L)

    def forward(self, x):
        return self.model(x)


class OracleUnetBlock(nn.Module):
    """ganslate/nn/generators/unet/unet2d.py:81-157 and unet3d.py:81-157 (the same block over 2-D / 3-D layers;
    norm_type 'instance': conv bias on, the outermost up-convolution keeps the ConvTranspose default = biased)."""

    def __init__(self, outer_nc, inner_nc, in_channels=None, submodule=None, outermost=False, innermost=False,
                 use_dropout=False, dims=2):
        super().__init__()
        self.outermost = outermost
        Conv, ConvT = (nn.Conv2d, nn.ConvTranspose2d) if dims == 2 else (nn.Conv3d, nn.ConvTranspose3d)
        norm = _norm2d if dims == 2 else nn.InstanceNorm3d
        in_channels = outer_nc if in_channels is None else in_channels
        downconv = Conv(in_channels, inner_nc, kernel_size=4, stride=2, padding=1, bias=True)
        downrelu, downnorm = nn.LeakyReLU(0.2), norm(inner_nc)
        uprelu, upnorm = nn.ReLU(), norm(outer_nc)
        if outermost:  # :123-127
            upconv = ConvT(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1)
            model = [downconv, submodule, uprelu, upconv, nn.Tanh()]
        elif innermost:  # :128-137
            upconv = ConvT(inner_nc, outer_nc, kernel_size=4, stride=2, padding=1, bias=True)
            model = [downrelu, downconv, uprelu, upconv, upnorm]
        else:  # :138-151
            upconv = ConvT(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1, bias=True)
            model = [downrelu, downconv, downnorm, submodule, uprelu, upconv, upnorm]
            if use_dropout:
                model.append(nn.Dropout(0.5))
        self.model = nn.Sequential(*model)

    def forward(self, x):  # :153-157
        return self.model(x) if self.outermost else torch.cat([x, self.model(x)], 1)


class OracleUnet2D(nn.Module):
    """ganslate/nn/generators/unet/unet2d.py:17-78 (dims=3: unet3d.py:17-78)"""

    def __init__(self, in_channels, out_channels, num_downs, ngf=64, use_dropout=False, dims=2):
        super().__init__()
        b = OracleUnetBlock(ngf * 8, ngf * 8, innermost=True, dims=dims)
        for _ in range(num_downs - 5):
            b = OracleUnetBlock(ngf * 8, ngf * 8, submodule=b, use_dropout=use_dropout, dims=dims)
        b = OracleUnetBlock(ngf * 4, ngf * 8, submodule=b, dims=dims)
        b = OracleUnetBlock(ngf * 2, ngf * 4, submodule=b, dims=dims)
        b = OracleUnetBlock(ngf, ngf * 2, submodule=b, dims=dims)
        self.model = OracleUnetBlock(out_channels, ngf, in_channels=in_channels, submodule=b, outermost=True, dims=dims)

    def forward(self, x):
        return self.model(x)


class OracleUnet3D(OracleUnet2D):

    def __init__(self, in_channels, out_channels, num_downs, ngf=64, use_dropout=False):
        super().__init__(in_channels, out_channels, num_downs, ngf, use_dropout, dims=3)


def init_weights(net, gain=0.02):
    """ganslate/nn/utils.py:13-36 with weight_init_type='normal': N(0, gain) on every Conv/Linear weight in
    module order, zero bias."""

    def f(m):
        name = type(m).__name__
        if hasattr(m, "weight") and ("Conv" in name or "Linear" in name):
            nn.init.normal_(m.weight.data, 0.0, gain)
            if getattr(m, "bias", None) is not None:
                nn.init.constant_(m.bias.data, 0.0)

    net.apply(f)
    return net


# --------------------------------------------------------------------------------------------- losses
def adversarial_lsgan(pred, target_is_real, real_label=1.0, fake_label=0.0):
    """ganslate/nn/losses/adversarial_loss.py:34-62 (gan_mode 'lsgan'): MSE against the expanded constant label."""
    t = torch.tensor(real_label if target_is_real else fake_label, dtype=pred.dtype, device=pred.device)
    return F.mse_loss(pred, t.expand_as(pred))


def ssim_distance(X, Y, data_range=1.0, win_size=11, win_sigma=1.5, K=(0.01, 0.03)):
    """ganslate/nn/losses/utils/ssim.py:22-99: separable Gaussian ("valid") moments, S1 * S2-style terms, mean of
    sqrt(relu(2 - S1 - S2)).  5-D inputs are viewed as (N*C) x D x H x W (:70-72): depth slices act as channels."""
    if X.ndim == 5:
        X, Y = X.reshape(-1, *X.shape[2:]), Y.reshape(-1, *Y.shape[2:])
    ch = X.shape[1]
    coords = torch.arange(win_size, dtype=X.dtype).float() - win_size // 2     # :36-41
    g = torch.exp(-(coords ** 2) / (2 * win_sigma ** 2))
    g = (g / g.sum()).to(X.dtype)
    win = g.view(1, 1, 1, -1).repeat(ch, 1, 1, 1)

    def blur(t):                                                                 # :44-49
        t = F.conv2d(t, win, groups=ch)
        return F.conv2d(t, win.transpose(2, 3), groups=ch)

    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2                  # :80-81
    mu1, mu2 = blur(X), blur(Y)
    s1, s2, s12 = blur(X * X) - mu1 ** 2, blur(Y * Y) - mu2 ** 2, blur(X * Y) - mu1 * mu2
    S1 = (2 * mu1 * mu2 + C1) / (mu1 ** 2 + mu2 ** 2 + C1)
    S2 = (2 * s12 + C2) / (s1 + s2 + C2)
    return torch.sqrt(torch.relu(2 - (S1 + S2))).mean()                          # :95-99


def cycle_loss(real, rec, proportion_ssim=0.0):
    """CycleLoss, ganslate/nn/losses/cyclegan_losses.py:60-91."""
    l1 = F.l1_loss(rec, real)
    if proportion_ssim > 0:
        return proportion_ssim * ssim_distance((rec + 1) / 2, (real + 1) / 2, 1.0) + (1 - proportion_ssim) * l1
    return l1


def cyclegan_losses(visuals, lambda_AB=10.0, lambda_BA=10.0, lambda_identity=0.0, proportion_ssim=0.0):
    """ganslate/nn/losses/cyclegan_losses.py:31-58,73-101 (proportion_ssim = 0 in every shipped YAML)."""
    out = {
        "cycle_A": lambda_AB * cycle_loss(visuals["real_A"], visuals["rec_A"], proportion_ssim),
        "cycle_B": lambda_BA * cycle_loss(visuals["real_B"], visuals["rec_B"], proportion_ssim),
    }
    if lambda_identity > 0:
        out["idt_B"] = lambda_AB * (F.l1_loss(visuals["idt_B"], visuals["real_B"]) * lambda_identity)
        out["idt_A"] = lambda_BA * (F.l1_loss(visuals["idt_A"], visuals["real_A"]) * lambda_identity)
    return out


class OracleImagePool:
    """ganslate/data/utils/image_pool.py:24-60"""

    def __init__(self, pool_size):
        self.pool_size, self.num_imgs, self.images = pool_size, 0, []

    def query(self, images):
        if self.pool_size == 0:
            return images
        ret = []
        for image in images:
            image = torch.unsqueeze(image.data, 0)
            if self.num_imgs < self.pool_size:
                self.num_imgs += 1
                self.images.append(image)
                ret.append(image)
            elif random.uniform(0, 1) > 0.5:
                i = random.randint(0, self.pool_size - 1)
                tmp = self.images[i].clone()
                self.images[i] = image
                ret.append(tmp)
            else:
                ret.append(image)
        return torch.cat(ret, 0)


# --------------------------------------------------------------------------------------------- CycleGAN step
def default_cyclegan_conf(**kw):
    c = dict(lambda_AB=10.0, lambda_BA=10.0, lambda_identity=0.0, lr_G=2e-4, lr_D=2e-4, beta1=0.5, beta2=0.999,
             pool_size=50, n_residual_blocks=9, ndf=64, n_layers=3, in_channels=3, out_channels=3)
    c.update(kw)
    return SimpleNamespace(**c)


class OracleCycleGAN:
    """One training iteration as ganslate/nn/gans/unpaired/cyclegan.py:92-214 runs it (mixed_precision False)."""

    def __init__(self, conf=None, seed=0, device="cpu"):
        self.conf = conf or default_cyclegan_conf()
        c = self.conf
        torch.manual_seed(seed)
        # construction + init order follows the dict order at cyclegan.py:52 / base.py:51-67
        self.networks = {}
        for name in ("G_AB", "G_BA", "D_B", "D_A"):
            if name.startswith("G"):
                net = OracleResnet2D(c.in_channels, c.out_channels, c.n_residual_blocks)
            else:
                net = OraclePatchGAN2D(c.in_channels, c.ndf, c.n_layers)
            self.networks[name] = init_weights(net).to(device)
        n = self.networks
        self.optimizers = {  # cyclegan.py:70-82
            "G": torch.optim.Adam(itertools.chain(n["G_AB"].parameters(), n["G_BA"].parameters()), lr=c.lr_G,
                                  betas=(c.beta1, c.beta2)),
            "D": torch.optim.Adam(itertools.chain(n["D_B"].parameters(), n["D_A"].parameters()), lr=c.lr_D,
                                  betas=(c.beta1, c.beta2)),
        }
        self.fake_A_pool, self.fake_B_pool = OracleImagePool(c.pool_size), OracleImagePool(c.pool_size)
        self.visuals, self.losses = {}, {}

    @staticmethod
    def _set_requires_grad(nets, flag):  # base.py:289-300
        for net in nets:
            for p in net.parameters():
                p.requires_grad = flag

    def forward(self):  # cyclegan.py:126-152
        n, v = self.networks, self.visuals
        v["fake_B"] = n["G_AB"](v["real_A"])
        v["rec_A"] = n["G_BA"](v["fake_B"])
        v["fake_A"] = n["G_BA"](v["real_B"])
        v["rec_B"] = n["G_AB"](v["fake_A"])
        v["idt_A"] = v["idt_B"] = None
        if self.conf.lambda_identity > 0:
            v["idt_B"] = n["G_AB"](v["real_B"])
            v["idt_A"] = n["G_BA"](v["real_A"])

    def backward_G(self):  # cyclegan.py:191-214
        n, v, c = self.networks, self.visuals, self.conf
        self.losses["G_AB"] = adversarial_lsgan(n["D_B"](v["fake_B"]), True)
        self.losses["G_BA"] = adversarial_lsgan(n["D_A"](v["fake_A"]), True)
        lg = cyclegan_losses(v, c.lambda_AB, c.lambda_BA, c.lambda_identity)
        self.losses.update(lg)
        (sum(lg.values()) + self.losses["G_AB"] + self.losses["G_BA"]).backward()

    def backward_D(self, name):  # cyclegan.py:154-189
        v = self.visuals
        if name == "D_B":
            real, fake = v["real_B"], self.fake_B_pool.query(v["fake_B"])
        else:
            real, fake = v["real_A"], self.fake_A_pool.query(v["fake_A"])
        pred_real = self.networks[name](real)
        pred_fake = self.networks[name](fake.detach())
        self.losses[name] = adversarial_lsgan(pred_real, True) + adversarial_lsgan(pred_fake, False)
        self.losses[name].backward()

    def optimize_parameters(self, real_A, real_B, step_optimizers=True):  # cyclegan.py:84-124
        self.visuals["real_A"], self.visuals["real_B"] = real_A, real_B
        ds = [self.networks["D_B"], self.networks["D_A"]]
        self.forward()
        self._set_requires_grad(ds, False)
        self.optimizers["G"].zero_grad(set_to_none=True)
        self.backward_G()
        if step_optimizers:
            self.optimizers["G"].step()
        self._set_requires_grad(ds, True)
        self.optimizers["D"].zero_grad(set_to_none=True)
        self.backward_D("D_B")
        self.backward_D("D_A")
        if step_optimizers:
            self.optimizers["D"].step()
        return {k: float(v.detach()) for k, v in self.losses.items() if v is not None}


def synthetic_batch(batch, channels=3, size=256, seed=1, device="cpu", width=None):
    """Inputs U(-1, 1) (images are normalised to [-1, 1] in the reference, utils/trackers/utils.py:81-82)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = width or size
    a = torch.rand((batch, channels, size, w), generator=g) * 2 - 1
    b = torch.rand((batch, channels, size, w), generator=g) * 2 - 1
    return a.to(device), b.to(device)


# --------------------------------------------------------------------------------------------- bf16 rounding points
# The fp32 oracle above answers "what does the reference compute".  The B200 path stores activations and feeds
# the tensor cores in bf16; a ReLU whose pre-activation is within one bf16 ulp of zero can flip, and every flip
# changes a gradient element by O(1), so element-wise gradient agreement with an fp32 run is bounded by
# sqrt(flip fraction) (~5 % per normalised layer) for ANY bf16 implementation.  To check the kernels themselves
# the same CPU restatement can be run with round-to-bf16 inserted at exactly the points where the B200 path
# rounds (DESIGN.md "Precision"): conv operands, conv outputs, fused norm/activation outputs, and the gradients
# wrt conv outputs.  Everything else (accumulation, statistics, losses, weight gradients, Adam) stays fp32.
# ROUND_BF16 = False turns every rounding point of forward_bf16_points into the identity: the same walk (and the same
# TRACE points) in plain fp32 -- the oracle of the fp32 validation mode's teacher-forced test.
ROUND_BF16 = True


class _RoundSTE(torch.autograd.Function):
    """forward: round to bf16; backward: identity."""

    @staticmethod
    def forward(ctx, x):
        if not ROUND_BF16:
            return x.view_as(x)
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


class _GradRound(torch.autograd.Function):
    """forward: identity; backward: round the gradient to bf16."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if not ROUND_BF16:
            return g
        return g.to(torch.bfloat16).to(torch.float32)


def _rb(x):
    return _RoundSTE.apply(x)


def _apply_act(act, y):
    if isinstance(act, nn.LeakyReLU):
        return F.leaky_relu(y, act.negative_slope)
    if isinstance(act, nn.ReLU):
        return F.relu(y)
    return torch.tanh(y)


def _flatten(mods):
    out = []
    for m in mods:
        if isinstance(m, nn.Sequential):
            out += _flatten(list(m))
        else:
            out.append(m)
    return out


# Noise floor of ANY bf16 implementation: with SUMMATION_VARIANT set, every convolution of forward_bf16_points adds
# its input channels in the reverse order -- mathematically the same network with the same rounding points, differing
# only in fp32 summation order (1e-7 relative).  Each bf16 rounding turns a difference d << ulp into sqrt(ulp * d), so
# after a few layers the two realisations are as far apart as two independent bf16 runs; the distance between them is
# the resolution at which an end-to-end comparison with the bf16-point oracle can detect anything
# (tests/test_cyclegan_gpu.py::test_cyclegan_step_vs_matched_oracle_noise_floor).
SUMMATION_VARIANT = False

TRACE = None  # debugging aid: list receiving ("raw"|"act", tensor) for every conv group of forward_bf16_points


TRACE_LIVE = False  # True: TRACE receives the live tensors (retain_grad) so that their gradients can be read too


def _trace(kind, t):
    if TRACE is not None:
        if TRACE_LIVE:
            if t.requires_grad:
                t.retain_grad()
            TRACE.append((kind, t))
        else:
            TRACE.append((kind, t.detach().clone()))


def _seq_bf16(mods, h, residual=None):
    """Walk [pad] conv [norm] [act] groups with the B200 path's rounding points. `h` is bf16-valued."""
    mods = _flatten(mods)
    i, n = 0, len(mods)
    while i < n:
        m = mods[i]
        if isinstance(m, (nn.ReflectionPad2d,)):
            h = m(h)
            i += 1
            continue
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Conv3d, nn.ConvTranspose3d)):
            w = _rb(m.weight)
            hh = h
            if SUMMATION_VARIANT:
                # the same sum in another order: input channels reversed in both operands (see SUMMATION_VARIANT)
                tr = isinstance(m, (nn.ConvTranspose2d, nn.ConvTranspose3d))
                hh, w = h.flip(1), w.flip(0 if tr else 1)
            if isinstance(m, nn.ConvTranspose2d):
                raw = F.conv_transpose2d(hh, w, m.bias, m.stride, m.padding, m.output_padding)
            elif isinstance(m, nn.ConvTranspose3d):
                raw = F.conv_transpose3d(hh, w, m.bias, m.stride, m.padding, m.output_padding)
            elif isinstance(m, nn.Conv3d):
                raw = F.conv3d(hh, w, m.bias, m.stride, m.padding)
            else:
                raw = F.conv2d(hh, w, m.bias, m.stride, m.padding)
            j = i + 1
            norm = j < n and isinstance(mods[j], (nn.InstanceNorm2d, nn.InstanceNorm3d))
            if norm:
                j += 1
            act = mods[j] if j < n and isinstance(mods[j], (nn.ReLU, nn.LeakyReLU, nn.Tanh)) else None
            if act is not None:
                j += 1
            last = j >= n
            if isinstance(act, nn.Tanh) and last:
                # generator output: bf16 pre-activation, tanh evaluated in fp32 on export
                raw = _GradRound.apply(_rb(raw))
                _trace("raw", raw)
                return torch.tanh(raw)
            followed_by_pad = j < n and (isinstance(mods[j], nn.ReflectionPad2d) or isinstance(mods[j], OracleResBlock))
            if norm or (last and residual is not None) or followed_by_pad:
                raw = _GradRound.apply(_rb(raw))                       # conv epilogue stores bf16
                _trace("raw", raw)
                y = F.instance_norm(raw, eps=1e-5) if norm else raw    # fp32 statistics
                if act is not None:
                    y = _apply_act(act, y)
                if last and residual is not None:
                    y = y + residual
                h = _rb(y)                                             # fused norm/act/residual kernel stores bf16
                _trace("act", h)
            else:
                pre = _GradRound.apply(raw)                            # activation in the conv epilogue: one rounding
                h = _rb(_apply_act(act, pre) if act is not None else pre)
                _trace("act", h)
            i = j
            continue
        if isinstance(m, OracleResBlock):
            h = _seq_bf16(list(m.conv_block), h, residual=h)
            i += 1
            continue
        raise NotImplementedError(type(m).__name__)
    return h


def forward_bf16_points(net, x):
    """Network forward with the B200 path's bf16 rounding points (network input rounded on entry)."""
    return _seq_bf16(list(net.model), _rb(x))


class OracleCycleGANBf16(OracleCycleGAN):
    """Same iteration as OracleCycleGAN, every network evaluated through forward_bf16_points."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        for net in self.networks.values():
            net.forward = (lambda x, _n=net: forward_bf16_points(_n, x))


# --------------------------------------------------------------------------------------------- Pix2Pix step
class OraclePix2Pix:
    """One iteration as ganslate/nn/gans/paired/pix2pix.py:76-152 runs it (Resnet2D generator, PatchGAN2D on
    cat[A, B]); lambda_pix2pix * L1 (pix2pix_losses.py:14-19)."""

    def __init__(self, lambda_pix2pix=30.0, n_residual_blocks=9, n_layers=4, seed=0, lr=2e-4, bf16_points=False,
                 unet=None):
        """unet: None (Resnet2D generator) or dict(num_downs, ngf) for the Unet2D generator of
        projects/cityscapes_label2photo/experiments/pix2pix.yaml (dropout off: parity needs determinism)."""
        torch.manual_seed(seed)
        self.lam = lambda_pix2pix
        gen = OracleResnet2D(3, 3, n_residual_blocks) if unet is None else OracleUnet2D(3, 3, unet["num_downs"],
                                                                                          ngf=unet["ngf"])
        self.networks = {"G": init_weights(gen),                                          # dict order pix2pix.py:42
                         "D": init_weights(OraclePatchGAN2D(6, 64, n_layers))}
        if bf16_points:
            for net in self.networks.values():
                net.forward = (lambda x, _n=net: forward_bf16_points(_n, x))
        self.optimizers = {k: torch.optim.Adam(self.networks[k].parameters(), lr=lr, betas=(0.5, 0.999)) for k in "GD"}
        self.visuals, self.losses = {}, {}

    def optimize_parameters(self, real_A, real_B, step_optimizers=True):
        G, D = self.networks["G"], self.networks["D"]
        fake_B = G(real_A)
        self.visuals = {"real_A": real_A, "real_B": real_B, "fake_B": fake_B}
        OracleCycleGAN._set_requires_grad([D], False)
        self.optimizers["G"].zero_grad(set_to_none=True)
        self.losses["G"] = adversarial_lsgan(D(torch.cat([real_A, fake_B], 1)), True)
        self.losses["pix2pix"] = self.lam * F.l1_loss(fake_B, real_B)
        (self.losses["G"] + self.losses["pix2pix"]).backward()
        if step_optimizers:
            self.optimizers["G"].step()
        OracleCycleGAN._set_requires_grad([D], True)
        self.optimizers["D"].zero_grad(set_to_none=True)
        pred_real = D(torch.cat([real_A, real_B], 1))
        pred_fake = D(torch.cat([real_A, fake_B.detach()], 1))
        self.losses["D"] = adversarial_lsgan(pred_real, True) + adversarial_lsgan(pred_fake, False)
        self.losses["D"].backward()
        if step_optimizers:
            self.optimizers["D"].step()
        return {k: float(v.detach()) for k, v in self.losses.items()}


# --------------------------------------------------------------------------------------------- CUT step
class OracleFeaturePatchMLP(nn.Module):
    """ganslate/nn/gans/unpaired/cut.py:229-294 (FeaturePatchMLP + LNorm)."""

    def __init__(self, channels_per_feature, num_patches=256, nc=256):
        super().__init__()
        self.num_patches = num_patches
        self.mlps = nn.ModuleList(
            [nn.Sequential(nn.Linear(c, nc), nn.ReLU(), nn.Linear(nc, nc)) for c in channels_per_feature])

    def forward(self, feats, patch_ids=None):
        out, ids = [], []
        for i, feat in enumerate(feats):
            feat = feat.permute(0, 2, 3, 1).flatten(1, 2)                       # :257-262 (2-D)
            pid = patch_ids[i] if patch_ids is not None else torch.randperm(feat.shape[1])[:self.num_patches]
            x = self.mlps[i](feat[:, pid, :].flatten(0, 1))                     # :270-275
            x = x / (x.pow(2).sum(1, keepdim=True).pow(0.5) + 1e-7)             # LNorm :290-294
            out.append(x)
            ids.append(pid)
        return out, ids


def patchnce_loss(feat_q, feat_k, batch_size, nce_T=0.07):
    """ganslate/nn/losses/cut_losses.py:14-43"""
    bs, dim = feat_q.shape[:2]
    feat_k = feat_k.detach()
    l_pos = torch.bmm(feat_q.view(bs, 1, -1), feat_k.view(bs, -1, 1)).view(bs, 1)
    q = feat_q.view(batch_size, -1, dim)
    k = feat_k.view(batch_size, -1, dim)
    n = q.size(1)
    l_neg = torch.bmm(q, k.transpose(2, 1))
    l_neg.masked_fill_(torch.eye(n, dtype=torch.bool)[None, :, :], -10.0)
    out = torch.cat((l_pos, l_neg.view(-1, n)), dim=1) / nce_T
    return F.cross_entropy(out, torch.zeros(out.size(0), dtype=torch.long), reduction="none")


def extract_features(x, net, layer_ids):
    """cut.py:297-312 -- iterate `net.encoder`, record after the listed indices (in-place ReLUs included)."""
    feats, feat = [], x
    for i, layer in enumerate(net.encoder):
        feat = layer(feat)
        if i in layer_ids:
            feats.append(feat)
    return feats


class OracleCUT:
    """One iteration as ganslate/nn/gans/unpaired/cut.py:113-226 runs it (Resnet2D + PatchGAN2D, lsgan)."""

    def __init__(self, n_residual_blocks=9, nce_layers=(0, 4, 8, 12, 16), num_patches=256, mlp_nc=256, batch_size=1,
                 lambda_adv=1.0, lambda_nce=1.0, lambda_nce_idt=0.5, nce_T=0.07, lr=2e-4, seed=0):
        torch.manual_seed(seed)
        self.nce_layers, self.batch_size, self.nce_T = tuple(nce_layers), batch_size, nce_T
        self.l_adv, self.l_nce, self.l_idt = lambda_adv, lambda_nce, lambda_nce_idt
        G = init_weights(OracleResnet2D(3, 3, n_residual_blocks))          # network dict order cut.py:72: G, D, mlp
        D = init_weights(OraclePatchGAN2D(3))
        with torch.no_grad():                                              # probe_network_channels cut.py:315-333
            chans = [f.shape[1] for f in extract_features(torch.zeros(1, 3, 64, 64), G, self.nce_layers)]
        mlp = init_weights(OracleFeaturePatchMLP(chans, num_patches, mlp_nc))
        self.networks = {"G": G, "D": D, "mlp": mlp}
        self.optimizers = {k: torch.optim.Adam(v.parameters(), lr=lr, betas=(0.5, 0.999)) for k, v in self.networks.items()}
        self.visuals, self.losses = {}, {}

    def _nce(self, source, target, patch_ids):
        G, mlp = self.networks["G"], self.networks["mlp"]
        sf = extract_features(source, G, self.nce_layers)
        tf = extract_features(target, G, self.nce_layers)
        sp, ids = mlp(sf, patch_ids)
        tp, _ = mlp(tf, ids)
        total = 0
        for t, s in zip(tp, sp):
            total = total + (patchnce_loss(t, s, self.batch_size, self.nce_T) * self.l_nce).mean()
        return total / len(self.nce_layers), ids

    def optimize_parameters(self, real_A, real_B, patch_ids=None, step_optimizers=True):
        G, D = self.networks["G"], self.networks["D"]
        fake_B = G(real_A)
        idt_B = G(real_B) if self.l_idt > 0 else None
        self.visuals = {"real_A": real_A, "real_B": real_B, "fake_B": fake_B, "idt_B": idt_B}
        OracleCycleGAN._set_requires_grad([D], True)
        self.optimizers["D"].zero_grad(set_to_none=True)
        self.losses["D"] = adversarial_lsgan(D(real_B), True) + adversarial_lsgan(D(fake_B.detach()), False)
        self.losses["D"].backward()
        if step_optimizers:
            self.optimizers["D"].step()
        OracleCycleGAN._set_requires_grad([D], False)
        self.optimizers["G"].zero_grad(set_to_none=True)
        self.optimizers["mlp"].zero_grad(set_to_none=True)
        self.losses["G"] = adversarial_lsgan(D(fake_B), True) * self.l_adv
        nce, ids = self._nce(real_A, fake_B, patch_ids)
        self.losses["NCE"] = nce
        total_nce = nce
        if self.l_idt > 0:
            nce_idt, _ = self._nce(real_B, idt_B, ids if patch_ids is None else patch_ids)
            self.losses["NCE_idt"] = self.l_idt * nce_idt
            total_nce = (1 - self.l_idt) * nce + self.losses["NCE_idt"]
        (self.losses["G"] + total_nce).backward()
        if step_optimizers:
            self.optimizers["G"].step()
            self.optimizers["mlp"].step()
        return {k: float(v.detach()) for k, v in self.losses.items()}, ids
