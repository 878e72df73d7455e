"""(a) Strided gathers as TMA boxes with traversal strides (stride-2 convolutions forward, transposed-convolution
data gradients, and their weight gradients) and (b) the pixel-window formulation of the generators' 7x7 input
convolution, against torch on the same inputs (tolerance 1e-2 max-relative per tensor, tests/gpu_bringup.py).
Each case also runs with the feature switched off as the control, and asserts which kernel served it."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

TMA, GATHER = 2, 1


def _run(knobs, expect_data=None, expect_wgrad=None, **case):
    import gpu_bringup
    from ganslate_b200 import _cabi
    lib = _cabi.lib()
    old = {k: lib.gb_debug_knob(k, v) for k, v in knobs.items()}
    lib.gb_debug_knob(15, 0)
    lib.gb_debug_knob(14, -1)
    try:
        ok = gpu_bringup.conv_case(**case)
        torch.cuda.synchronize()
        data_path, wgrad_path = lib.gb_debug_knob(15, 0), lib.gb_debug_knob(14, 0)
    finally:
        for k, v in old.items():
            lib.gb_debug_knob(k, v)
    assert ok
    if expect_data is not None:
        assert data_path == expect_data, f"last gb_conv_data call was served by kernel path {data_path}"
    if expect_wgrad is not None:
        assert wgrad_path == expect_wgrad, f"last gb_conv_wgrad call was served by variant {wgrad_path}"


STRIDED = [
    dict(name="3x3 s2 p1 64->128 64x64", cin=64, cout=128, k=3, s=2, p=1, H=64, W=64, N=2),
    dict(name="3x3 s2 p1 128->256 30x46 ragged", cin=128, cout=256, k=3, s=2, p=1, H=30, W=46, N=3),
    dict(name="4x4 s2 p1 64->128 33x33", cin=64, cout=128, k=4, s=2, p=1, H=33, W=33),
    dict(name="convT 3x3 s2 p1 op1 128->64 32x32", cin=128, cout=64, k=3, s=2, p=1, H=32, W=32, N=2, transposed=True,
         op_pad=1),
    dict(name="3d 2x2x2 s2 64->128 8x16x16", cin=64, cout=128, k=2, s=2, p=0, H=16, W=16, D=8),
]


@pytest.mark.parametrize("case", STRIDED, ids=[c["name"] for c in STRIDED])
def test_strided_gather_through_tma(case):
    # conv: forward is the strided gather (last data call = its data gradient, parity classes, TMA already);
    # transposed conv: the data gradient is the strided gather.  Weight gradient: TMA variant 1.
    _run({}, expect_data=TMA, expect_wgrad=1, **case)
    _run({0: 2}, expect_wgrad=0, **dict(case, name=case["name"] + " (control: cp.async gather)"))


WINDOW = [
    dict(name="7x7 reflect3 3->64 64x64", cin=3, cout=64, k=7, s=1, p=0, H=64, W=64, reflect=3),
    dict(name="7x7 reflect3 3->64 37x53 N=3", cin=3, cout=64, k=7, s=1, p=0, H=37, W=53, N=3, reflect=3),
    dict(name="7x7 p0 1->128 40x40", cin=1, cout=128, k=7, s=1, p=0, H=40, W=40, N=2),
    dict(name="5x5 p0 8->64 24x31", cin=8, cout=64, k=5, s=1, p=0, H=24, W=31),
]


@pytest.mark.parametrize("case", WINDOW, ids=[c["name"] for c in WINDOW])
def test_pixel_window_convolution(case):
    from ganslate_b200 import _cabi, ops
    assert _cabi.lib().gb_tma_window_supported() == 1, "driver rejected the overlapping-stride tensor map"
    _run({}, expect_wgrad=1, **case)          # window: TMA-fed weight gradient
    old = ops.WINDOW_CONV
    ops.WINDOW_CONV = False
    try:
        _run({}, expect_wgrad=0, **dict(case, name=case["name"] + " (control: tap formulation)"))
    finally:
        ops.WINDOW_CONV = old


def test_window_forward_is_tma_fed():
    from ganslate_b200 import _cabi, ops
    lib = _cabi.lib()
    op = ops.ConvOp(3, 64, (1, 7, 7), (1, 1, 1), (0, 0, 0))
    assert op.window
    w = torch.randn(64, 3, 7, 7, device="cuda") * 0.1
    x = torch.randn(2, 1, 38, 38, 8, device="cuda").to(torch.bfloat16)
    x[..., 3:] = 0
    lib.gb_debug_knob(15, 0)
    y = op.run_fwd(ops.make_view(x), "cuda", w, None)
    torch.cuda.synchronize()
    assert lib.gb_debug_knob(15, 0) == TMA
    ref = torch.nn.functional.conv2d(x[:, 0, :, :, :3].permute(0, 3, 1, 2).float(), w.to(torch.bfloat16).float())
    got = y[:, 0].permute(0, 3, 1, 2).float()
    assert ((got - ref).abs().max() / ref.abs().max()).item() < 1e-2


BWD_WINDOW = [
    dict(name="7x7 reflect3 64->3 tanh 64x64", cin=64, cout=3, k=7, s=1, p=0, H=64, W=64, reflect=3, act="tanh"),
    dict(name="7x7 reflect3 64->3 37x53 N=3", cin=64, cout=3, k=7, s=1, p=0, H=37, W=53, N=3, reflect=3),
    dict(name="7x7 p0 128->1 40x40", cin=128, cout=1, k=7, s=1, p=0, H=40, W=40, N=2),
    dict(name="5x5 p1 64->8 24x31", cin=64, cout=8, k=5, s=1, p=1, H=24, W=31),
]


@pytest.mark.parametrize("case", BWD_WINDOW, ids=[c["name"] for c in BWD_WINDOW])
def test_gradient_side_pixel_windows(case):
    """Output layers with <= 8 channels: data gradient and operand-swapped weight gradient through 7-pixel windows of
    the zero-bordered dOut (ops.ConvOp.bwd_window) -- both TMA-fed; control = the tap formulation on the gather path."""
    from ganslate_b200 import ops
    old_bw, ops.BWD_WINDOW_CONV = ops.BWD_WINDOW_CONV, True
    try:
        _run({}, expect_data=TMA, expect_wgrad=1, **case)
    finally:
        ops.BWD_WINDOW_CONV = old_bw
    old = ops.WINDOW_CONV
    ops.WINDOW_CONV = False
    try:
        _run({}, expect_wgrad=0, **dict(case, name=case["name"] + " (control: tap formulation)"))
    finally:
        ops.WINDOW_CONV = old
