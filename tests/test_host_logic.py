"""Host-side logic that needs no GPU: implicit-GEMM geometry (class / tap decomposition and weight packing
strides) checked against torch's CPU convolutions, the config tree, and the plug-in builders."""
import itertools

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from ganslate_b200 import ops
from ganslate_b200.configs.utils import Conf, init_config
from ganslate_b200.presets import cyclegan_resnet2d


def emulate_data_gemm(spec, x, wmat_fn, out_ext, n_out):
    """Evaluate a DataSpec literally as include/ganslate_b200.h defines gb_conv_data:
    out[q*out_mul+off][n] = sum_{t,c} in[q*in_mul + d_t][c] * W[n][t][c], zero outside the input."""
    N, C, D, H, W = x.shape
    out = np.zeros((N, n_out) + tuple(out_ext), dtype=np.float64)
    xn = x.numpy().astype(np.float64)
    for cls in spec.classes:
        off = cls["off"]
        q_ext = [max(0, -(-(out_ext[d] - off[d]) // spec.out_mul[d])) for d in range(3)]
        for q in itertools.product(*[range(e) for e in q_ext]):
            o = tuple(q[d] * spec.out_mul[d] + off[d] for d in range(3))
            for tl, (tap, tid) in enumerate(zip(cls["taps"], cls["tap_ids"])):
                i = tuple(q[d] * spec.in_mul[d] + tap[d] for d in range(3))
                if all(0 <= i[d] < (D, H, W)[d] for d in range(3)):
                    out[:, :, o[0], o[1], o[2]] += xn[:, :, i[0], i[1], i[2]] @ wmat_fn(tid).T
    return out


CASES = [
    dict(k=(1, 3, 3), s=(1, 1, 1), p=(0, 1, 1), transposed=False),
    dict(k=(1, 3, 3), s=(1, 2, 2), p=(0, 1, 1), transposed=False),
    dict(k=(1, 4, 4), s=(1, 2, 2), p=(0, 1, 1), transposed=False),
    dict(k=(1, 7, 7), s=(1, 1, 1), p=(0, 0, 0), transposed=False),
    dict(k=(1, 3, 3), s=(1, 2, 2), p=(0, 1, 1), transposed=True, op=(0, 1, 1)),
    dict(k=(1, 4, 4), s=(1, 2, 2), p=(0, 1, 1), transposed=True, op=(0, 0, 0)),
    dict(k=(2, 2, 2), s=(2, 2, 2), p=(0, 0, 0), transposed=False),
    dict(k=(2, 2, 2), s=(2, 2, 2), p=(0, 0, 0), transposed=True, op=(0, 0, 0)),
    dict(k=(3, 3, 3), s=(1, 1, 1), p=(1, 1, 1), transposed=False),
]


@pytest.mark.parametrize("case", CASES)
def test_specs_reproduce_torch_convolutions(case):
    torch.manual_seed(0)
    cin, cout = 3, 5
    k, s, p = case["k"], case["s"], case["p"]
    tr = case["transposed"]
    op = ops.ConvOp(cin, cout, k, s, p, transposed=tr, output_padding=case.get("op", (0, 0, 0)))
    x = torch.randn(2, cin, 1, 9, 8) if not tr else torch.randn(2, cin, 1, 4, 3)
    if k[0] > 1:
        x = torch.randn(2, cin, 4, 6, 6)
    w = torch.randn((cin, cout) + k) if tr else torch.randn((cout, cin) + k)
    x.requires_grad_(True)
    if tr:
        y = F.conv_transpose3d(x, w, None, s, p, case.get("op", (0, 0, 0)))
    else:
        y = F.conv3d(x, w, None, s, p)
    assert tuple(y.shape[2:]) == op.out_extent(tuple(x.shape[2:]))
    wn = w.numpy().astype(np.float64).reshape(w.shape[0], w.shape[1], -1)
    # forward: rows = cout, k-channels = cin
    fwd_w = (lambda tid: wn[:, :, tid].T) if tr else (lambda tid: wn[:, :, tid])
    got = emulate_data_gemm(op.fwd, x.detach(), fwd_w, tuple(y.shape[2:]), cout)
    np.testing.assert_allclose(got, y.detach().numpy(), rtol=1e-4, atol=1e-4)
    # data gradient: rows = cin, k-channels = cout
    g = torch.randn_like(y)
    (dx,) = torch.autograd.grad(y, x, g)
    dg_w = (lambda tid: wn[:, :, tid]) if tr else (lambda tid: wn[:, :, tid].T)
    got = emulate_data_gemm(op.dgrad, g, dg_w, tuple(x.shape[2:]), cin)
    np.testing.assert_allclose(got, dx.numpy(), rtol=1e-4, atol=1e-4)
    # packing strides address the same elements as the emulation's matrices
    sn, sc, st = op.fwd_strides
    flat = w.contiguous().view(-1)
    for n, c, t in [(0, 0, 0), (cout - 1, cin - 1, op.T - 1), (1, 2, op.T // 2)]:
        assert flat[n * sn + c * sc + t * st].item() == pytest.approx(fwd_w(t)[n, c])
    sn, sc, st = op.dgrad_strides
    for n, c, t in [(0, 0, 0), (cin - 1, cout - 1, op.T - 1)]:
        assert flat[n * sn + c * sc + t * st].item() == pytest.approx(dg_w(t)[n, c])


def emulate_wgrad(op, x, g):
    """Evaluate ConvOp.wgrad_plan() literally as include/ganslate_b200.h defines gb_conv_wgrad and gb_unpack_wgrad:
    dw[r][t*Cg + c] = sum_q plain[q][r] * gathered[q*mul + d_t][c] (zero outside), then
    dst[r*dsr + c*dsc + t*dst_t] = dw[r][t*chans_pad + c]."""
    plan = op.wgrad_plan()
    xn, gn = x.numpy().astype(np.float64), g.numpy().astype(np.float64)
    plain, gathered = (xn, gn) if plan["plain_is_input"] else (gn, xn)
    u = plan["unpack"]
    Cg_pad = u["chans_pad"]
    dw = np.zeros((plan["rows_pad"], plan["kpad"]))
    N, _, D, H, W = plain.shape
    ext_g = gathered.shape[2:]
    for q in itertools.product(range(D), range(H), range(W)):
        for t, tap in enumerate(plan["taps"]):
            i = tuple(q[d] * plan["mul"][d] + tap[d] for d in range(3))
            if all(0 <= i[d] < ext_g[d] for d in range(3)):
                # [rows] x [gathered channels], summed over the batch
                blk = plain[:, :, q[0], q[1], q[2]].T @ gathered[:, :, i[0], i[1], i[2]]
                dw[:blk.shape[0], t * Cg_pad:t * Cg_pad + blk.shape[1]] += blk
    dst = np.zeros(op.cout * op.cin * op.T)
    for r in range(u["rows"]):
        for c in range(u["chans"]):
            for t in range(u["ntaps"]):
                dst[r * u["dsr"] + c * u["dsc"] + t * u["dst_t"]] = dw[r, t * Cg_pad + c]
    return dst


@pytest.mark.parametrize("cin,cout,k,s,p,tr", [
    (3, 5, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),
    (3, 5, (1, 4, 4), (1, 2, 2), (0, 1, 1), False),
    (40, 3, (1, 3, 3), (1, 1, 1), (0, 0, 0), False),    # few output channels: operand swap
    (136, 1, (1, 4, 4), (1, 1, 1), (0, 1, 1), False),   # PatchGAN's last layer shape class: operand swap
    (5, 3, (1, 3, 3), (1, 2, 2), (0, 1, 1), True),
])
def test_wgrad_plan_reproduces_torch_weight_gradient(cin, cout, k, s, p, tr):
    torch.manual_seed(0)
    op = ops.ConvOp(cin, cout, k, s, p, transposed=tr, output_padding=(0, 1, 1) if tr else (0, 0, 0))
    x = torch.randn(2, cin, 1, 6, 5)
    w = (torch.randn((cin, cout) + k) if tr else torch.randn((cout, cin) + k)).requires_grad_(True)
    y = F.conv_transpose3d(x, w, None, s, p, (0, 1, 1)) if tr else F.conv3d(x, w, None, s, p)
    g = torch.randn_like(y)
    (dw,) = torch.autograd.grad(y, w, g)
    got = emulate_wgrad(op, x, g)
    np.testing.assert_allclose(got, dw.numpy().reshape(-1), rtol=1e-4, atol=1e-4)
    assert op.wg_swap == (cout <= 3 and not tr and s == (1, 1, 1) and cin >= 40)


def test_spec_padding_and_limits():
    op = ops.ConvOp(3, 64, (1, 7, 7), (1, 1, 1), (0, 0, 0))
    assert op.cin_pad == 8 and op.fwd.kpads == [448] and op.wg_kpad == 448
    op = ops.ConvOp(256, 256, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    assert len(op.dgrad.classes) == 4 and sorted(len(c["taps"]) for c in op.dgrad.classes) == [1, 2, 2, 4]
    with pytest.raises(ValueError):
        ops.ConvOp(8, 8, (6, 6, 6), (1, 1, 1), (0, 0, 0))
    assert op.flops((1, 128, 128), 1) == 2.0 * 64 * 64 * 256 * 256 * 9


def test_config_tree_and_target_defaults():
    conf = cyclegan_resnet2d(batch_size=2)
    assert conf.mode == "train" and conf["train"].batch_size == 2
    gan = conf.train.gan
    assert gan.norm_type == "instance" and gan.weight_init_gain == 0.02 and gan.pool_size == 50
    assert tuple(gan.generator.in_out_channels.BA) == (3, 3)          # BA <- AB (configs/base.py:30)
    assert gan.discriminator.in_channels.A == 3                       # A <- B (configs/base.py:42)
    assert gan.discriminator.kernel_size == (4, 4) and gan.discriminator.ndf == 64
    assert gan.optimizer.beta1 == 0.5 and gan.optimizer.lambda_AB == 10.0
    c2 = init_config(conf.to_dict(), overrides=["train.gan.optimizer.lr_G=0.001", "train.batch_size=4"])
    assert c2.train.gan.optimizer.lr_G == 0.001 and c2.train.batch_size == 4
    assert isinstance(conf.train, Conf) and "gan" in conf.train and dict(gan.discriminator)["n_layers"] == 3


def test_builders_construct_networks_in_reference_order():
    from ganslate_b200.utils.builders import build_D, build_G
    conf = cyclegan_resnet2d()
    torch.manual_seed(0)
    g = build_G(conf, "AB", "cpu")
    d = build_D(conf, "A", "cpu")
    assert sum(p.numel() for p in g.parameters()) == 11378179
    assert sum(p.numel() for p in d.parameters()) == 2764737
    assert float(g.model[1].bias.abs().sum()) == 0.0 and 0.015 < float(g.model[1].weight.std()) < 0.025


def test_scheduler_rule_matches_reference_formula():
    from ganslate_b200.nn.utils import get_scheduler
    conf = cyclegan_resnet2d(n_iters=10, n_iters_decay=10)
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1.0)
    sch = get_scheduler(opt, conf)
    lrs = []
    for _ in range(22):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    # lr_l = 1 - max(0, iter + 1 - n_iters) / (n_iters_decay + 1)   (ganslate/nn/utils.py:91-97)
    want = [1.0 - max(0, i + 1 - 10) / 11.0 for i in range(22)]
    assert lrs == pytest.approx(want)


def test_pixel_window_formulation_reproduces_torch():
    """ConvOp.window (k x 7 convolution over 3 -> 8 channels read as 7-pixel windows of a 64-channel view): the
    forward GEMM with the packed pseudo-tap weights and the weight gradient with its per-K-block unpack items,
    evaluated literally as include/ganslate_b200.h defines them, against torch."""
    torch.manual_seed(0)
    old = ops.WINDOW_CONV
    ops.WINDOW_CONV = "force"
    try:
        op = ops.ConvOp(3, 64, (1, 7, 7), (1, 1, 1), (0, 0, 0))
    finally:
        ops.WINDOW_CONV = old
    assert op.window and op.fwd.kpads == [7 * 64] and op.wg_kpad == 7 * 64 and len(op.fwd.classes[0]["taps"]) == 7
    N, H, W, kw, cin, cout, T = 2, 11, 12, 7, 3, 64, 49
    x = torch.randn(N, cin, 1, H, W, requires_grad=True)
    w = torch.randn(cout, cin, 1, 7, 7, requires_grad=True)
    y = F.conv3d(x, w)
    g = torch.randn_like(y)
    y.backward(g)
    # channels-last, 8 physical channels, flat; window view: C = 64 of which kw*8 exist, pixel stride 8
    cl = np.zeros((N, H, W, 8))
    cl[..., :cin] = x.detach().numpy()[:, :, 0].transpose(0, 2, 3, 1)
    flat = np.concatenate([cl.reshape(-1), np.full(64, np.nan)])  # reads past kw*8 must never happen

    def window(n, yy, xx):  # 64 values: kw*8 from memory, the rest zero-filled by the TMA unit
        base = ((n * H + yy) * W + xx) * 8
        return np.concatenate([flat[base:base + kw * 8], np.zeros(64 - kw * 8)])

    Wout = W - kw + 1
    # packed weights exactly as pack_class() computes them
    wn = w.detach().numpy().reshape(cout, cin, T)
    wp = np.zeros((cout, op.fwd.kpads[0]))
    for k in range(op.fwd.kpads[0]):
        tl, c = divmod(k, 8)
        tid = op.win_pack_ids[tl] if tl < len(op.win_pack_ids) else -1
        if tid >= 0 and c < cin:
            wp[:, k] = wn[:, c, tid]
    taps = op.fwd.classes[0]["taps"]
    got = np.zeros((N, cout, H - 6, Wout))
    for n in range(N):
        for yy in range(H - 6):
            for xx in range(Wout):
                a = np.concatenate([window(n, yy + t[1], xx + t[2]) for t in taps])
                got[n, :, yy, xx] = wp @ a
    np.testing.assert_allclose(got, y.detach().numpy()[:, :, 0], rtol=1e-4, atol=1e-4)
    # weight gradient: dw[r][b*64 + k] = sum_q g[q][r] * window[q + tap_b][k], then the unpack items
    gn = g.numpy()[:, :, 0]
    dw = np.zeros((op.wg_rows_pad, op.wg_kpad))
    for n in range(N):
        for yy in range(H - 6):
            for xx in range(Wout):
                a = np.concatenate([window(n, yy + t[1], xx + t[2]) for t in op.wg_taps])
                dw[:cout] += np.outer(gn[n, :, yy, xx], a)
    dst = np.zeros(cout * cin * T)
    for it in op.window_unpack_items():
        for r in range(it["rows"]):
            for c in range(it["chans"]):
                for t in range(it["ntaps"]):
                    dst[it["dst_off"] + r * it["dsr"] + c * it["dsc"] + t * it["dst_t"]] = \
                        dw[r, it["ws_off"] + t * it["chans_pad"] + c]
    np.testing.assert_allclose(dst.reshape(cout, cin, 7, 7), w.grad.numpy()[:, :, 0], rtol=1e-4, atol=1e-4)


def test_gradient_side_pixel_windows_reproduce_torch():
    """ConvOp.bwd_window (7x7 convolution 64 -> 3 channels): data gradient and operand-swapped weight gradient
    gathered as 7-pixel windows of the zero-bordered dOut rows, evaluated literally, against torch autograd."""
    torch.manual_seed(1)
    old = ops.WINDOW_CONV
    ops.WINDOW_CONV = "force"
    try:
        op = ops.ConvOp(64, 3, (1, 7, 7), (1, 1, 1), (0, 0, 0))
    finally:
        ops.WINDOW_CONV = old
    assert op.bwd_window and not op.window and op.wg_swap
    N, Hin, Win, cin, cout, kw, T, BL = 2, 10, 12, 64, 3, 7, 49, ops.BWD_BORDER
    Ho, Wo = Hin - 6, Win - 6
    x = torch.randn(N, cin, 1, Hin, Win, requires_grad=True)
    w = torch.randn(cout, cin, 1, 7, 7, requires_grad=True)
    y = F.conv3d(x, w)
    g = torch.randn_like(y)
    y.backward(g)
    # zero-bordered channels-last dOut: (N, Ho, Wo + 2*BL, 8), flat, NaN sentinel behind it
    gw = np.zeros((N, Ho, Wo + 2 * BL, 8))
    gw[:, :, BL:BL + Wo, :cout] = g.numpy()[:, :, 0].transpose(0, 2, 3, 1)
    Wb = Wo + 2 * BL
    flat = np.concatenate([gw.reshape(-1), np.full(64, np.nan)])

    def window(n, yy, xs):  # window at row yy, start column xs of the buffer; rows outside the tensor read as zero
        if not (0 <= yy < Ho):
            return np.zeros(64)
        assert 0 <= xs <= Wb - kw, xs
        base = ((n * Ho + yy) * Wb + xs) * 8
        return np.concatenate([flat[base:base + kw * 8], np.zeros(64 - kw * 8)])

    # packed data-gradient weights as pack_class() builds them from bw_pack_ids and dgrad_strides
    sn, sc, st = op.dgrad_strides
    wflat = w.detach().numpy().reshape(-1)
    kp = op.dgrad.kpads[0]
    wp = np.zeros((cin, kp))
    for k in range(kp):
        tl, c = divmod(k, 8)
        tid = op.bw_pack_ids[tl] if tl < len(op.bw_pack_ids) else -1
        if tid >= 0 and c < cout:
            wp[:, k] = [wflat[n * sn + c * sc + tid * st] for n in range(cin)]
    taps = op.dgrad.classes[0]["taps"]
    dx = np.zeros((N, cin, Hin, Win))
    for n in range(N):
        for yy in range(Hin):
            for xx in range(Win):
                a = np.concatenate([window(n, yy + t[1], xx + t[2]) for t in taps])
                dx[n, :, yy, xx] = wp @ a
    np.testing.assert_allclose(dx, x.grad.numpy()[:, :, 0], rtol=1e-4, atol=1e-4)
    # swapped weight gradient: dw[ci][blk*64 + j*8 + co] = sum_q' x[q'][ci] * window[q' + tap_blk][j*8 + co]
    xn = x.detach().numpy()[:, :, 0]
    dw = np.zeros((op.wg_rows_pad, op.wg_kpad))
    for n in range(N):
        for yy in range(Hin):
            for xx in range(Win):
                a = np.concatenate([window(n, yy + t[1], xx + t[2]) for t in op.wg_taps])
                dw[:cin] += np.outer(xn[n, :, yy, xx], a)
    dst = np.zeros(cout * cin * T)
    for it in op.window_unpack_items():
        for r in range(it["rows"]):
            for c in range(it["chans"]):
                for t in range(it["ntaps"]):
                    dst[it["dst_off"] + r * it["dsr"] + c * it["dsc"] + t * it["dst_t"]] = \
                        dw[r, it["ws_off"] + t * it["chans_pad"] + c]
    np.testing.assert_allclose(dst.reshape(cout, cin, 7, 7), w.grad.numpy()[:, :, 0], rtol=1e-4, atol=1e-4)


def test_replicate_pad_index_math_matches_torch():
    """The index math of csrc/pad.cu restated in numpy: forward dst[z,y,x] = src[clamp(z-pz), clamp(y-py), clamp(x-px)];
    backward dsrc[i] = sum of ddst over readers(i) = [0, p] for i = 0, [n-1+p, n-1+2p] for i = n-1, {i+p} otherwise
    (per axis) -- against torch's ReplicationPad3d and its autograd, incl. extent-1 axes."""
    def readers(i, n, p):
        lo = 0 if i == 0 else i + p
        hi = n - 1 + 2 * p if i == n - 1 else i + p
        return range(lo, hi + 1)

    for (D, H, W), (pz, py, px) in [((3, 4, 5), (1, 1, 1)), ((1, 6, 2), (2, 3, 1)), ((2, 1, 4), (0, 2, 2))]:
        torch.manual_seed(D * 10 + W)
        x = torch.randn(1, 1, D, H, W, dtype=torch.float64, requires_grad=True)
        y = F.pad(x, (px, px, py, py, pz, pz), mode="replicate")
        g = torch.randn_like(y)
        y.backward(g)
        xn, gn = x.detach().numpy()[0, 0], g.numpy()[0, 0]
        fwd = np.empty((D + 2 * pz, H + 2 * py, W + 2 * px))
        for z in range(D + 2 * pz):
            for yy in range(H + 2 * py):
                for xx in range(W + 2 * px):
                    fwd[z, yy, xx] = xn[min(max(z - pz, 0), D - 1), min(max(yy - py, 0), H - 1), min(max(xx - px, 0), W - 1)]
        np.testing.assert_array_equal(fwd, y.detach().numpy()[0, 0])
        bwd = np.zeros((D, H, W))
        for z in range(D):
            for yy in range(H):
                for xx in range(W):
                    bwd[z, yy, xx] = sum(gn[a, b, c] for a in readers(z, D, pz) for b in readers(yy, H, py)
                                         for c in readers(xx, W, px))
        np.testing.assert_allclose(bwd, x.grad.numpy()[0, 0], rtol=1e-12, atol=1e-12)


def test_device_image_pool_makes_the_reference_pools_decisions():
    """DeviceImagePool (one preallocated tensor, gather + scatter per query) returns, image by image, what the
    reference's list-based ImagePool (ganslate/data/utils/image_pool.py:24-60) returns for the same `random` stream --
    including a slot hit twice inside one batch and the fill phase ending mid-batch."""
    import random
    import torch
    from ganslate_b200.data.utils.image_pool import DeviceImagePool, ImagePool
    g = torch.Generator().manual_seed(0)
    for pool_size, batch in ((5, 3), (50, 8), (4, 8), (0, 2)):
        ref, dev = ImagePool(pool_size), DeviceImagePool(pool_size)
        for it in range(40):
            x = torch.randn(batch, 2, 3, 3, generator=g)
            random.seed(1000 + it)
            a = ref.query(x.clone())
            random.seed(1000 + it)
            b = dev.query(x.clone())
            assert torch.equal(a, b), (pool_size, batch, it)
        if pool_size:
            assert dev.num_imgs == ref.num_imgs
            for k in range(ref.num_imgs):
                assert torch.equal(ref.images[k][0], dev.store[k]), (pool_size, k)


def test_fused_adam_state_dict_has_one_step_per_parameter():
    """the shared device `step` scalar of a FusedAdam group must not reach a checkpoint -- torch.optim.Adam
    restored from it would advance `step` once per parameter.  state_dict() emits per-parameter CPU scalars."""
    import torch
    from ganslate_b200.optim.fused_adam import FusedAdam
    ps = [torch.nn.Parameter(torch.randn(3)) for _ in range(4)]
    opt = FusedAdam(ps, lr=1e-3)
    shared = torch.tensor(5.0)
    for p in ps:                       # the state FusedAdam.step() leaves behind (one tensor shared by the group)
        opt.state[p] = {"exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p), "step": shared}
    sd = opt.state_dict()
    steps = [st["step"] for st in sd["state"].values()]
    assert all(float(s) == 5.0 for s in steps) and len({s.data_ptr() for s in steps}) == 4
    ref = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=1e-3)
    ref.load_state_dict(sd)
    for p in ref.param_groups[0]["params"]:
        p.grad = torch.ones_like(p)
    ref.step()
    assert all(float(st["step"]) == 6.0 for st in ref.state.values())
