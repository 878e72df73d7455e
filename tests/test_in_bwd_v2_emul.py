"""Host run of the second-generation fused InstanceNorm backward (ganslate_b200/csrc/instnorm_v2_core.h, opt-in on
the GPU through gb_debug_knob(22, 1 | 2)): the per-thread body that nvcc compiles into `in_bwd_v2_kernel` is compiled
with g++ (tests/emul/in_bwd_v2_emul.cpp) and every thread of every block is run on the CPU, against the pointer-level
restatement of gb_in_bwd (tests/fake_cabi.py, which follows include/ganslate_b200.h).  Covers what can go wrong
without a GPU in the loop: row-segment walking with ragged block ranges, the reflection fold (rows, columns,
corners; border 1 and 2), channel groups that do not fill the block, 3-D row-linear views, interior views of
bordered buffers, the residual-gradient accumulation happening exactly once, every dx element written exactly once,
the FMA-form algebra and the bf16 packing.  The block reduction, atomics and grid barrier are device-only and are
what the GPU tests of knob 22 (tests/test_ops_gpu.py cases under GB_KNOBS=22=1) add."""
import ctypes as C
import os
import subprocess

import pytest
import torch

from ganslate_b200 import _cabi
from ganslate_b200._cabi import ACT_LEAKY, ACT_NONE, ACT_PRELU, ACT_RELU

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = tmp_path_factory.mktemp("emul") / "in_bwd_emul.so"
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", os.path.join(HERE, "emul", "in_bwd_v2_emul.cpp"),
           os.path.join(HERE, "emul", "in_bwd_v3_emul.cpp"), "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    assert res.returncode == 0, res.stderr
    lib = C.CDLL(str(out))
    for fn in (lib.in_bwd_v2_emulate, lib.in_bwd_v3_emulate):
        fn.argtypes = [C.POINTER(_cabi.InBwdParams), C.c_int, C.c_int, C.c_float, C.POINTER(C.c_int)]
        fn.restype = C.c_int
    lib.in_fwd_v2_emulate.argtypes = [C.POINTER(_cabi.InFwdParams), C.c_int, C.c_int, C.c_float]
    lib.in_fwd_v2_emulate.restype = C.c_int
    return lib


def _view(t, pad, N, D, H, W, Cc, c0=0):
    """gb_view of the interior of a (N, D, H + 2 pad, W + 2 pad, Ctot) channels-last tensor, channels c0 .. c0 + Cc."""
    v = _cabi.View()
    sn, sz, sy, sx, _ = t.stride()
    v.ptr = t.data_ptr() + (pad * sy + pad * sx + c0) * t.element_size()
    v.sn, v.sz, v.sy, v.sx = sn, sz, sy, sx
    v.N, v.D, v.H, v.W, v.C, v.pad = N, D, H, W, Cc, pad
    return v


CASES = [
    # N, D, H,  W,  C,   gpad, act,       res,   cap, U, x_border, c_slice
    (2, 1, 12, 10, 64, 1, ACT_RELU, False, 7, 4, 0, False),     # ragged block ranges, 8 slots... of 32
    (1, 1, 9, 7, 8, 0, ACT_LEAKY, False, 5, 2, 0, False),       # one channel group: 256 pixels per step > image
    (3, 1, 8, 8, 256, 1, ACT_NONE, True, 10, 3, 1, False),      # resblock tail: residual gradient, x inside a border
    (2, 3, 5, 6, 16, 0, ACT_RELU, False, 9, 4, 0, False),       # 3-D, row-linear
    (1, 1, 11, 13, 24, 1, ACT_LEAKY, False, 4, 4, 0, False),    # 3 channel groups: 85 slots, one idle thread
    (2, 1, 8, 9, 32, 2, ACT_RELU, True, 6, 2, 2, False),        # border 2: two mirrored rows / columns per side
    (1, 1, 16, 16, 64, 1, ACT_RELU, False, 1, 4, 0, False),     # a single block per image walks 16 whole rows
    (2, 1, 6, 20, 40, 1, ACT_NONE, False, 64, 3, 0, True),      # more blocks than rows; channel slice of a wider buffer
    (5, 1, 7, 7, 16, 1, ACT_RELU, False, 3, 4, 0, False),       # more images than co-resident blocks (two-launch form)
]


def _ids(cases):
    return [f"N{c[0]}D{c[1]}H{c[2]}W{c[3]}C{c[4]}p{c[5]}a{c[6]}r{int(c[7])}cap{c[8]}U{c[9]}" for c in cases]


@pytest.mark.parametrize("case", CASES, ids=_ids(CASES))
def test_v2_thread_body_matches_the_abi_restatement(emul, case):
    def run(p, cap, U, ns):
        grid = (C.c_int * 2)()
        assert emul.in_bwd_v2_emulate(C.byref(p), cap, U, ns, grid) == 0
        assert grid[0] >= 1 and grid[0] * grid[1] >= case[1] * case[2] * case[3]
    _check(case, run)


def _check(case, run):
    import fake_cabi
    N, D, H, W, Cc, gpad, act, res, cap, U, xb, c_slice = case[:12]
    gen = case[12] if len(case) > 12 else {}
    torch.manual_seed(sum(int(v) for v in case[:12]) % 1000)
    Ctot = Cc + 16 if c_slice else Cc
    c0 = 8 if c_slice else 0
    x_t = (torch.randn(N, D, H + 2 * xb, W + 2 * xb, Ctot) * 1.5 + 0.7).to(torch.bfloat16)
    dy_t = torch.randn(N, D, H + 2 * gpad, W + 2 * gpad, Ctot)
    xv = _view(x_t, xb, N, D, H, W, Cc, c0)
    xi = x_t[:, :, xb:xb + H, xb:xb + W, c0:c0 + Cc].float()
    stats = torch.stack([xi.sum(dim=(1, 2, 3)), (xi * xi).sum(dim=(1, 2, 3))], dim=-1).contiguous()
    slope = 0.2 if act == ACT_LEAKY else 0.0
    ns = {ACT_NONE: 1.0, ACT_RELU: 0.0, ACT_LEAKY: slope, ACT_PRELU: 0.0}[act]
    sum0 = torch.randn(N, D, H, W, Ctot) if res else None
    prelu = (torch.rand(Cc) * 0.5 + 0.05) if act == ACT_PRELU else None
    res_t = (torch.randn(N, D, H, W, Ctot) * 0.8).to(torch.bfloat16) if gen.get("rba") else None

    def params(dx_t, dysum_t, bstats, dbias, dprelu=None):
        p = _cabi.InBwdParams()
        p.x = xv
        p.dy_b = _view(dy_t, gpad, N, D, H, W, Cc, c0)
        p.dx = _view(dx_t, xb, N, D, H, W, Cc, c0)
        if dysum_t is not None:
            p.dy_sum = _view(dysum_t, 0, N, D, H, W, Cc, c0)
            p.dy_sum_acc = 1
        p.stats, p.bstats, p.dbias = stats.data_ptr(), bstats.data_ptr(), dbias.data_ptr()
        p.eps, p.act, p.act_slope = 1e-5, act, slope
        if prelu is not None:
            p.prelu = prelu.data_ptr()
            if dprelu is not None:
                p.dprelu = dprelu.data_ptr()
        if res_t is not None:
            p.res = _view(res_t, 0, N, D, H, W, Cc, c0)
            p.res_before_act = 1
        p.out_scale = gen.get("oscale", 0.0)
        return p

    outs = []
    for which in ("ref", "v2"):
        dx_t = torch.full(x_t.shape, float("nan")).to(torch.bfloat16)
        dysum_t = sum0.clone() if res else None
        bstats = torch.zeros(N * Cc * 2 + 4)
        dbias = torch.zeros(Cc)
        dprelu = torch.zeros(Cc) if gen.get("dprelu") else None
        p = params(dx_t, dysum_t, bstats, dbias, dprelu)
        if which == "ref":
            assert fake_cabi.FakeLib().gb_in_bwd(p, None) == 0
        else:
            run(p, cap, U, ns)
        outs.append((dx_t, dysum_t, dbias, dprelu))
    (dx_r, sum_r, db_r, dp_r), (dx_v, sum_v, db_v, dp_v) = outs
    if dp_r is not None:
        assert torch.allclose(dp_v, dp_r, rtol=1e-3, atol=1e-3 * max(1.0, dp_r.abs().max().item()))
    # dx: interior written everywhere (no NaN sentinel left), nothing outside the interior / channel slice touched
    inner = dx_v[:, :, xb:xb + H, xb:xb + W, c0:c0 + Cc].float()
    assert not torch.isnan(inner).any()
    mask = torch.ones_like(dx_v, dtype=torch.bool)
    mask[:, :, xb:xb + H, xb:xb + W, c0:c0 + Cc] = False
    assert torch.isnan(dx_v.float()[mask]).all()
    ref = dx_r[:, :, xb:xb + H, xb:xb + W, c0:c0 + Cc].float()
    # same algebra in a different association: agreement to bf16 rounding of values of the gradient's scale
    scale = ref.abs().max().item()
    assert (inner - ref).abs().max().item() <= 2.0 ** -7 * scale
    assert (inner - ref).abs().mean().item() <= 2.0 ** -10 * scale
    assert torch.allclose(db_v, db_r, rtol=1e-3, atol=1e-3 * max(1.0, db_r.abs().max().item()))
    if res:
        assert torch.allclose(sum_v, sum_r, rtol=1e-5, atol=1e-5)
        assert torch.equal(sum_v[..., :c0], sum0[..., :c0]) and torch.equal(sum_v[..., c0 + Cc:], sum0[..., c0 + Cc:])


# the general form of the second generation (V-Net layers): U = 12 selects it in the emulator
CASES_GEN = [
    (2, 3, 6, 7, 16, 0, ACT_PRELU, False, 9, 12, 0, False, dict(dprelu=True)),                   # PReLU + its gradient
    (1, 2, 9, 8, 32, 0, ACT_PRELU, True, 5, 12, 0, False, dict(dprelu=True, rba=True)),          # V-Net: act(IN(x) + res)
    (2, 1, 8, 9, 24, 1, ACT_RELU, True, 6, 12, 1, False, dict(rba=True)),                        # masked residual gradient, 2-D fold
    (1, 4, 5, 5, 16, 0, ACT_PRELU, True, 7, 12, 0, True, dict(oscale=-1.0, dprelu=True)),        # inverse coupling: -act(IN(x)) + res
    (3, 1, 7, 6, 8, 0, ACT_NONE, False, 4, 12, 0, False, dict(oscale=0.5)),
    (2, 1, 12, 10, 64, 1, ACT_LEAKY, True, 7, 12, 0, False, dict()),                             # the plain cases through the general form
]


@pytest.mark.parametrize("case", CASES_GEN, ids=_ids(CASES_GEN))
def test_v2_general_form_matches_the_abi_restatement(emul, case):
    def run(p, cap, U, ns):
        grid = (C.c_int * 2)()
        assert emul.in_bwd_v2_emulate(C.byref(p), cap, U, ns, grid) == 0
    _check(case, run)


# the on-chip kernel (instnorm_v3_core.h): clusters of K CTAs own (image, 32 channels); `cap` is the SM count here
CASES_V3 = [
    (2, 1, 12, 10, 64, 1, ACT_RELU, False, 7, 4, 0, False),      # one CTA per cluster, two steps, ragged
    (3, 1, 8, 8, 256, 1, ACT_NONE, True, 10, 2, 1, False),       # residual gradient, x inside a border, one step
    (2, 3, 5, 6, 32, 0, ACT_RELU, False, 9, 4, 0, False),        # 3-D, row-linear
    (1, 1, 64, 64, 32, 1, ACT_RELU, False, 148, 4, 0, False),    # residual-block map: K grows to 8 to fill the SMs
    (1, 1, 90, 91, 32, 1, ACT_LEAKY, False, 4, 4, 0, False),     # 8190 pixels: K = 8 at full stash capacity (16 steps)
    (1, 1, 33, 31, 64, 2, ACT_RELU, True, 1, 2, 2, False),       # 1023 pixels in one CTA: 16 steps, last one ragged; border 2
    (2, 1, 20, 13, 32, 1, ACT_NONE, False, 300, 4, 0, True),     # channel slice of a wider buffer, K = 2
    (1, 1, 47, 45, 96, 1, ACT_LEAKY, True, 2, 4, 1, False),      # three channel groups, K = 4 with a ragged last range
    (8, 1, 64, 64, 32, 1, ACT_RELU, True, 2, 14, 0, False),      # plan mode 2: K = 8, half-size stash (8 steps)
]


@pytest.mark.parametrize("case", CASES_V3, ids=_ids(CASES_V3))
def test_v3_on_chip_thread_body_matches_the_abi_restatement(emul, case):
    P = case[1] * case[2] * case[3]

    def run(p, sms, U, ns):
        geom = (C.c_int * 3)()
        assert emul.in_bwd_v3_emulate(C.byref(p), sms, U, ns, geom) == 0
        K, ppc, steps = geom[0], geom[1], geom[2]
        assert 1 <= K <= 8 and K * ppc >= P and (K - 1) * ppc < P and steps <= 16 and steps * 64 >= ppc
        assert U < 10 or steps <= 8  # plan mode 2: half-size stash
    _check(case, run)


def test_v3_plan_limits(emul):
    """Maps that do not fit a cluster's shared memory, or channel counts that are not whole groups, are declined."""
    x = torch.zeros(1, 1, 4, 4, 32, dtype=torch.bfloat16)
    geom = (C.c_int * 3)()
    for (H, W, Cc, ok) in [(64, 64, 256, True), (128, 128, 128, False), (91, 90, 32, True), (91, 91, 32, False), (8, 8, 24, False)]:
        p = _cabi.InBwdParams()
        p.x = _view(x, 0, 1, 1, H, W, Cc)
        p.x.ptr = None  # declined (or accepted) before any memory is touched: plan() only
        if ok:
            continue  # accepted shapes are exercised above; here only the refusals
        assert emul.in_bwd_v3_emulate(C.byref(p), 148, 4, 0.0, geom) == -1


def test_v2_plan_covers_every_pixel_once():
    """gbv2::plan restated: equal pixel ranges per block, the last one ragged, never more blocks than `cap`."""
    for N, P, cap in [(8, 4096, 296), (8, 65536, 296), (1, 900, 296), (3, 17, 296), (300, 64, 296), (2, 1, 10)]:
        nb = max(cap // N, 1)
        ppb = max((P + nb - 1) // nb, 1)
        nblocks = (P + ppb - 1) // ppb
        assert nblocks <= nb and (nblocks - 1) * ppb < P <= nblocks * ppb


# forward (gbv2::fwd_pass): N, D, H, W, C, y border, act, residual, cap, U (14 = general form), x border, options
CASES_FWD = [
    (2, 1, 12, 10, 64, 1, ACT_RELU, False, 7, 4, 0, {}),
    (3, 1, 8, 8, 256, 1, ACT_NONE, True, 10, 2, 1, {}),                       # resblock tail: + residual, border for the next pad
    (1, 1, 11, 13, 24, 3, ACT_LEAKY, False, 4, 4, 0, {}),                     # border 3 (the 7x7 output layer's input)
    (2, 3, 5, 6, 16, 0, ACT_RELU, False, 9, 4, 0, {}),                        # 3-D
    (1, 1, 9, 7, 8, 0, ACT_RELU, False, 5, 2, 0, dict(no_norm=True)),         # activation only (stats == NULL)
    (2, 2, 6, 7, 16, 0, ACT_PRELU, False, 9, 14, 0, {}),                      # V-Net: PReLU
    (1, 2, 9, 8, 32, 0, ACT_PRELU, True, 5, 14, 0, dict(rba=True)),           # V-Net: act(IN(x) + res)
    (1, 4, 5, 5, 16, 0, ACT_PRELU, True, 7, 14, 0, dict(oscale=-1.0)),        # inverse coupling: res - act(IN(x))
    (2, 1, 8, 9, 24, 2, ACT_RELU, True, 6, 14, 0, dict(rba=True)),            # general form with a 2-D border
]


@pytest.mark.parametrize("case", CASES_FWD, ids=[f"N{c[0]}D{c[1]}H{c[2]}W{c[3]}C{c[4]}p{c[5]}a{c[6]}r{int(c[7])}U{c[9]}" for c in CASES_FWD])
def test_v2_forward_thread_body_matches_the_abi_restatement(emul, case):
    import fake_cabi
    N, D, H, W, Cc, ypad, act, res, cap, U, xb, opt = case
    torch.manual_seed(sum(int(v) for v in case[:11]))
    x_t = (torch.randn(N, D, H + 2 * xb, W + 2 * xb, Cc) * 1.5 + 0.7).to(torch.bfloat16)
    xv = _view(x_t, xb, N, D, H, W, Cc)
    xi = x_t[:, :, xb:xb + H, xb:xb + W].float()
    stats = torch.stack([xi.sum(dim=(1, 2, 3)), (xi * xi).sum(dim=(1, 2, 3))], dim=-1).contiguous()
    slope = 0.2 if act == ACT_LEAKY else 0.0
    ns = {ACT_NONE: 1.0, ACT_RELU: 0.0, ACT_LEAKY: slope, ACT_PRELU: 0.0}[act]
    prelu = (torch.rand(Cc) * 0.5 + 0.05) if act == ACT_PRELU else None
    res_t = (torch.randn(N, D, H, W, Cc) * 0.8).to(torch.bfloat16) if res else None
    outs = []
    for which in ("ref", "v2"):
        y_t = torch.full((N, D, H + 2 * ypad, W + 2 * ypad, Cc), float("nan")).to(torch.bfloat16)
        p = _cabi.InFwdParams()
        p.x, p.y = xv, _view(y_t, ypad, N, D, H, W, Cc)
        if res_t is not None:
            p.res = _view(res_t, 0, N, D, H, W, Cc)
            p.res_before_act = 1 if opt.get("rba") else 0
        p.stats = None if opt.get("no_norm") else stats.data_ptr()
        p.prelu = prelu.data_ptr() if prelu is not None else None
        p.eps, p.act, p.act_slope, p.out_scale = 1e-5, act, slope, opt.get("oscale", 0.0)
        if which == "ref":
            assert fake_cabi.FakeLib().gb_in_fwd(p, None) == 0
        else:
            assert emul.in_fwd_v2_emulate(C.byref(p), cap, U, ns) == 0
        outs.append(y_t.float())
    ref, got = outs
    assert not torch.isnan(got).any()           # interior and the whole reflection border written
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= 2.0 ** -7 * scale     # one bf16 rounding of a differently associated value
    assert (got - ref).abs().mean().item() <= 2.0 ** -10 * scale
