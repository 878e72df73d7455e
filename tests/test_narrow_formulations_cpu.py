"""The two reformulations behind the narrow-channel kernels, restated literally in torch on the CPU and compared with
torch's own convolution / weight gradient.  They pin the INDEX MATH (which pixel meets which tap where, what the zero
fill at the image border and at tile seams must do) independently of any GPU:

* csrc/igemm_xsplit.cu -- the dx taps of a kernel row as column blocks of one GEMM over un-shifted activation windows,
  `out[p] = sum_j P_j[p + dx_j]`, evaluated tile by tile (8 rows x 16 halo columns, 16 - (kw - 1) outputs per row);
* csrc/igemm_wgrad_narrow.cu -- `x` read at the right-most tap against `dOut` shifted by j = 0 .. kw - 1 pixels,
  `D[(dy, ci)][(j, co)] = dW[(dz, dy, dx_max - j)][ci][co]`, summed over 16 x 8 patches whose grid starts kw - 1 pixels
  left of the image so that every (x pixel, dOut pixel) pair meets exactly once."""
import pytest
import torch


def _fetch(x, z, y0, x0, h, w):
    """x: (C, D, H, W); the h x w window at depth z starting at (y0, x0), zero outside the tensor (TMA out-of-bounds fill)."""
    C, D, H, W = x.shape
    out = torch.zeros(C, h, w, dtype=x.dtype)
    if not 0 <= z < D:
        return out
    ya, yb = max(y0, 0), min(y0 + h, H)
    xa, xb = max(x0, 0), min(x0 + w, W)
    if ya < yb and xa < xb:
        out[:, ya - y0:yb - y0, xa - x0:xb - x0] = x[:, z, ya:yb, xa:xb]
    return out


@pytest.mark.parametrize("k,shape", [(5, (3, 20, 27)), (3, (4, 9, 33)), (5, (2, 8, 12))])
def test_xsplit_tiles_equal_the_convolution(k, shape):
    torch.manual_seed(0)
    cin, cout = 3, 4
    D, H, W = shape
    pad = k // 2
    x = torch.randn(cin, D, H, W, dtype=torch.float64)
    wgt = torch.randn(cout, cin, k, k, k, dtype=torch.float64)
    ref = torch.nn.functional.conv3d(x[None], wgt, padding=pad)[0]
    TY, HW = 8, 16
    xw = HW - (k - 1)
    dmin = -pad                                      # taps r - padding, r = 0 .. k - 1
    out = torch.zeros_like(ref)
    for z0 in range(D):
        for y0 in range(0, H, TY):
            for x0 in range(0, W, xw):
                # accumulator tile: rows (yl, x'), column blocks j; one "MMA" per (dz, dy) over the un-shifted window
                P = torch.zeros(k, cout, TY, HW, dtype=torch.float64)
                for dzi in range(k):
                    halo = _fetch(x, z0 + dmin + dzi, y0 + dmin, x0 + dmin, TY + k - 1, HW)
                    for dyi in range(k):
                        win = halo[:, dyi:dyi + TY, :]                           # contiguous 8 x 16 window
                        for j in range(k):                                      # column block j <-> dx = dmin + j
                            P[j] += torch.einsum("oc,cyx->oyx", wgt[:, :, dzi, dyi, j], win)
                # epilogue: out[w] = sum_j P_j[w + j] for the xw output columns of the tile row
                for wv in range(xw):
                    if x0 + wv >= W:
                        break
                    acc = sum(P[j][:, :, wv + j] for j in range(k))             # (cout, TY)
                    rows = min(TY, H - y0)
                    out[:, z0, y0:y0 + rows, x0 + wv] = acc[:, :rows]
    assert torch.allclose(out, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("k,shape", [(5, (3, 20, 27)), (3, (2, 17, 9)), (5, (1, 16, 8))])
def test_shifted_dout_patches_equal_the_weight_gradient(k, shape):
    torch.manual_seed(1)
    cin, cout = 3, 2
    D, H, W = shape
    pad = k // 2
    x = torch.randn(cin, D, H, W, dtype=torch.float64, requires_grad=False)
    wgt = torch.zeros(cout, cin, k, k, k, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(cout, D, H, W, dtype=torch.float64)
    torch.nn.functional.conv3d(x[None], wgt, padding=pad).backward(dy[None])
    ref = wgt.grad                                   # (cout, cin, kz, ky, kx)
    TW, TH, PITCH = 8, 16, 16
    dmin, dmax = -pad, k - 1 - pad
    got = torch.zeros_like(ref)
    ntx = -(-(W + k - 1) // TW)
    for dzi in range(k):                             # one CTA column per depth offset
        for z0 in range(D):
            for y0 in range(0, H, TH):
                for tx in range(ntx):
                    x0 = tx * TW - (k - 1)           # the patch grid starts kw - 1 pixels left of the image
                    # x box: pixel i <-> x0 + dmax + i, rows from y0 + dmin; dOut box: pixel i <-> x0 + i
                    xa = _fetch(x, z0 + dmin + dzi, y0 + dmin, x0 + dmax, TH + k - 1, PITCH)
                    ga = _fetch(dy, z0, y0, x0, TH, PITCH)
                    for r in range(TH):              # reduction over the patch's pixels p' = (r, i), i < 8
                        for dyi in range(k):         # M atoms: halo rows
                            xs = xa[:, r + dyi, :TW]                            # (cin, 8)
                            for j in range(k):       # N atoms: dOut shifted by j pixels  ->  dx = dmax - j
                                gs = ga[:, r, j:j + TW]                         # (cout, 8)
                                got[:, :, dzi, dyi, (dmax - j) - dmin] += torch.einsum("op,cp->oc", gs, xs)
    assert torch.allclose(got, ref, rtol=1e-10, atol=1e-10)
