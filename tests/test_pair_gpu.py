"""The two-patch halo kernel (csrc/igemm_pair.cu) and the two-row-tile weight-gradient variant (igemm_wgrad.cu,
RT = 2) against torch on the same inputs, for every tile width / halo pitch / stage configuration, forced through
the bring-up knobs so that small test shapes reach the kernels the batch-8 workload uses.

Tolerance: max |ours - ref| / max |ref| <= 1e-2 per tensor (tests/gpu_bringup.py::conv_case), as for every other
convolution kernel."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

PAIR, RT2 = 4, 2


def _run(knobs, expect_data=None, expect_wgrad=None, **case):
    import gpu_bringup
    from ganslate_b200 import _cabi
    lib = _cabi.lib()
    old = {k: lib.gb_debug_knob(k, v) for k, v in knobs.items()}
    lib.gb_debug_knob(15, 0)
    lib.gb_debug_knob(14, 0)
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


RESBLOCK = dict(name="3x3 reflect1 256->256 64x64 N=2", cin=256, cout=256, k=3, s=1, p=0, H=64, W=64, N=2, reflect=1)


@pytest.mark.parametrize("bn", [64, 128, 256])
@pytest.mark.parametrize("pitch16", [0, 1])
def test_pair_resblock_conv(bn, pitch16):
    _run({9: 2, 10: pitch16, 11: bn, 12: 2}, expect_data=PAIR, expect_wgrad=RT2, **RESBLOCK)


@pytest.mark.parametrize("a_stages", [1, 2])
def test_pair_activation_stage_counts(a_stages):
    _run({9: 2, 11: 256, 13: a_stages}, expect_data=PAIR, **RESBLOCK)


def test_pair_odd_patch_count_and_ragged_image():
    # 40 x 24 image: 3 x 3 = 9 patches (the last CTA has one), rows / columns beyond the image inside the patches
    _run({9: 2}, expect_data=PAIR, name="3x3 p1 64->64 40x24", cin=64, cout=64, k=3, s=1, p=1, H=40, W=24)
    _run({9: 2}, expect_data=PAIR, name="3x3 p1 128->64 19x23 N=3", cin=128, cout=64, k=3, s=1, p=1, H=19, W=23, N=3)


def test_pair_partial_column_tile():
    # 72 output channels: second column tile is partial (weight rows past the matrix are zero-filled by TMA);
    # the data gradient (72 input channels) falls back to the other kernels
    _run({9: 2}, name="3x3 p1 64->72 24x24", cin=64, cout=72, k=3, s=1, p=1, H=24, W=24, N=2)


def test_pair_patchgan_4x4_stride1():
    _run({9: 2, 11: 128, 12: 2}, expect_data=PAIR, expect_wgrad=RT2, name="4x4 s1 p1 256->512 32x32", cin=256, cout=512, k=4,
         s=1, p=1, H=32, W=32)
    _run({9: 2, 11: 256}, expect_data=PAIR, name="4x4 s1 p1 256->512 32x32 bn256", cin=256, cout=512, k=4, s=1, p=1,
         H=32, W=32)


def test_pair_3d_tap_groups():
    _run({9: 2}, expect_data=PAIR, name="3d 3x3x3 p1 64->64 4x16x16", cin=64, cout=64, k=3, s=1, p=1, H=16, W=16, D=4)


def test_default_heuristics_on_full_batch():
    # no knobs, batch 8 residual-block layer: forward = one wave of pair CTAs, data gradient (66x66 padded domain,
    # last gb_conv_data call of the case) = per-tap kernel with 11x11 patches, weight gradient = one row tile
    case = dict(RESBLOCK, name="3x3 reflect1 256->256 64x64 N=8", N=8)
    _run({}, expect_data=2, expect_wgrad=1, **case)


def test_non_power_of_two_patches():
    # per-tap TMA kernel on domains whose best patch is not a power of two (66x66 -> 11x11, 36x20 -> 18x7 ...),
    # against the power-of-two restriction (knob 0 = 1) as the control
    for knobs in ({9: 1}, {9: 1, 0: 1}):
        _run(knobs, expect_data=2, **dict(RESBLOCK, name=f"resblock 11x11 patches {knobs}"))
        _run(knobs, expect_data=2, name=f"3x3 p1 64->64 36x20 N=2 {knobs}", cin=64, cout=64, k=3, s=1, p=1, H=36, W=20, N=2)
        _run(knobs, expect_data=2, name=f"3x3 p1 128->128 13x50 {knobs}", cin=128, cout=128, k=3, s=1, p=1, H=13, W=50)
