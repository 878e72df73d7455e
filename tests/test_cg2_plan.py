"""Work decomposition of the persistent convolution kernels (csrc/igemm_cg2.cu) replayed on the host:
gb_debug_cg2_plan runs the SAME decode_item / tile_of code as the kernels (compiled __host__ __device__).  Every output
pixel of every parity class and every column block must be produced by exactly one valid (item, CTA rank), both CTAs
of a pair must agree on `any`, and invalid tiles must point at loadable coordinates."""
import ctypes as C
import itertools

import numpy as np
import pytest

from ganslate_b200 import _cabi, ops
from ganslate_b200._cabi import ConvParams, View


def _view(N, D, H, W, Cc, esz=2):
    v = View()
    v.ptr = 0x10000  # never dereferenced: the plan is host arithmetic only
    v.sn, v.sz, v.sy, v.sx = D * H * W * Cc, H * W * Cc, W * Cc, Cc
    v.N, v.D, v.H, v.W, v.C, v.pad = N, D, H, W, Cc, 0
    return v


CASES = [
    # (name, cin, cout, kernel, stride, padding, transposed, which, N, in extents)
    ("3x3 p1 fwd 256->256 64x64 N=8", 256, 256, (1, 3, 3), (1, 1, 1), (0, 1, 1), False, "fwd", 8, (1, 64, 64)),
    ("3x3 p0 dgrad onto 66x66 N=3", 256, 256, (1, 3, 3), (1, 1, 1), (0, 0, 0), False, "dgrad", 3, (1, 66, 66)),
    ("3x3 s2 p1 fwd 64->128 30x46 ragged", 64, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1), False, "fwd", 3, (1, 30, 46)),
    ("3x3 s2 p1 dgrad (4 parity classes) 128->64", 64, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1), False, "dgrad", 2, (1, 31, 33)),
    ("convT 3x3 s2 p1 op1 fwd 128->64 (4 classes)", 128, 64, (1, 3, 3), (1, 2, 2), (0, 1, 1), True, "fwd", 2, (1, 16, 16)),
    ("4x4 s1 p1 fwd 256->512 (2 column blocks)", 256, 512, (1, 4, 4), (1, 1, 1), (0, 1, 1), False, "fwd", 1, (1, 32, 32)),
    ("3d 3x3x3 p1 fwd 64->72 5x9x7 odd tiles", 64, 72, (3, 3, 3), (1, 1, 1), (1, 1, 1), False, "fwd", 1, (5, 9, 7)),
    ("1x1 fwd 64->64 one tile", 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), False, "fwd", 1, (1, 8, 8)),
]


@pytest.mark.parametrize("mode", [1, 2], ids=["cta-pair", "single-cta"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_every_output_pixel_is_produced_exactly_once(case, mode):
    name, cin, cout, k, s, pd, tr, which, N, ext = case
    lib = _cabi.lib()
    op = ops.ConvOp(cin, cout, k, s, pd, transposed=tr, output_padding=(0, 1, 1) if tr else (0, 0, 0))
    p = ConvParams()
    spec = op.fwd if which == "fwd" else op.dgrad
    op._fill(spec, p)
    if which == "fwd":
        out_ext = op.out_extent(ext)
        p.inp, p.out = _view(N, *ext, op.cin_pad), _view(N, *out_ext, op.cout_pad)
        p.ncols, p.npad = op.cout, op.fwd_rows_pad
    else:  # data gradient: gathers dOut (extents = forward output of `ext`), writes the input-shaped gradient
        dy_ext = op.out_extent(ext)
        out_ext = ext
        p.inp, p.out = _view(N, *dy_ext, op.cout_pad), _view(N, *ext, op.cin_pad, 4)
        p.ncols, p.npad = op.cin, op.dgrad_rows_pad
    info = (C.c_int32 * 6)()
    n = lib.gb_debug_cg2_plan(C.byref(p), mode, info, None, 0)
    if n < 0:
        pytest.skip("launch does not take the persistent path (class matrices of different K)")
    tw, th, bn, nitems, ranks, ntiles = list(info)
    assert ranks == (2 if mode == 1 else 1) and tw * th <= 128 and nitems > 0
    buf = (C.c_int32 * (nitems * ranks * 8))()
    assert lib.gb_debug_cg2_plan(C.byref(p), mode, info, buf, len(buf)) == nitems
    plan = np.frombuffer(buf, dtype=np.int32).reshape(nitems, ranks, 8)
    nb = -(-p.ncols // bn)
    cover = np.zeros((N, nb) + tuple(out_ext), dtype=np.int32)
    out_mul = list(p.out_mul)
    for i in range(nitems):
        anys = set(int(plan[i, r, 7]) for r in range(ranks))
        assert len(anys) == 1, "the CTAs of a pair disagree on whether the item exists"
        for r in range(ranks):
            cls, n0, x0, y0, z0, nn, valid, any_ = [int(v) for v in plan[i, r]]
            assert 0 <= cls < p.nclass and n0 % bn == 0 and 0 <= n0 < nb * bn
            off = list(p.cls[cls].off)
            q = [max(0, -(-(out_ext[d] - off[d]) // out_mul[d])) for d in range(3)]
            if not valid:
                assert (x0, y0, z0, nn) == (0, 0, 0, 0)  # loadable dummy coordinates
                continue
            assert any_ == 1 and 0 <= nn < N and 0 <= z0 < q[0] and 0 <= y0 < q[1] and 0 <= x0 < q[2]
            assert x0 % tw == 0 and y0 % th == 0
            for h, w in itertools.product(range(th), range(tw)):
                qy, qx = y0 + h, x0 + w
                if qy < q[1] and qx < q[2]:
                    cover[nn, n0 // bn, z0 * out_mul[0] + off[0], qy * out_mul[1] + off[1], qx * out_mul[2] + off[2]] += 1
    assert cover.min() == 1 and cover.max() == 1, (cover.min(), cover.max())
