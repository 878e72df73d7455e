"""The halo-reuse kernel for 16 / 32-channel inputs (csrc/igemm_halo_narrow.cu: SWIZZLE_32B / SWIZZLE_64B pixel rows read
through shifted K-major descriptors, weight boxes of 64 / C taps) against torch on the same inputs: the V-Net 5x5x5 layer
shapes, ragged images, 2-D windows, every column-tile width, forward and data gradient -- and the matching weight
gradient (csrc/igemm_wgrad_narrow.cu: one halo box per depth offset, MN-major descriptors whose atoms are one pixel apart).

Tolerance: max |ours - ref| / max |ref| <= 1e-2 per tensor (tests/gpu_bringup.py::conv_case), as for every other
convolution kernel."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

NARROW = 8      # knob 15 read-back: igemm_halo_narrow
XSPLIT = 9      # knob 15 read-back: igemm_xsplit (takes the layers with <= 32 output rows first)
WG_NARROW = 3   # knob 14 read-back: igemm_wgrad_narrow


def _case(expect_last="auto", knobs=None, expect_wgrad=None, **case):
    import gpu_bringup
    from ganslate_b200 import _cabi
    lib = _cabi.lib()
    old = {k: lib.gb_debug_knob(k, v) for k, v in (knobs or {}).items()}
    lib.gb_debug_knob(15, 0)
    lib.gb_debug_knob(14, 0)
    try:
        ok = gpu_bringup.conv_case(**case)
        torch.cuda.synchronize()
        path, wpath = lib.gb_debug_knob(15, 0), lib.gb_debug_knob(14, 0)
    finally:
        for k, v in old.items():
            lib.gb_debug_knob(k, v)
    assert ok
    if expect_last == "auto":   # the data gradient reads cout channels and writes cin: x-split takes 16 -> 16 by default
        expect_last = XSPLIT if (case["cin"] == 16 and case["cout"] == 16) else NARROW
    if expect_last is not None:
        assert path == expect_last, f"last gb_conv_data call (the data gradient) was served by kernel path {path}"
    if expect_wgrad is not None:
        assert wpath == expect_wgrad, f"the weight gradient was served by variant {wpath}"


@pytest.mark.parametrize("cin,cout", [(32, 32), (16, 16), (32, 16), (16, 32)])
def test_vnet_k5_layers(cin, cout):
    _case(expect_wgrad=WG_NARROW, name=f"3d k5 p2 {cin}->{cout} 6x32x24", cin=cin, cout=cout, k=5, s=1, p=2, H=32, W=24, D=6)


def test_ragged_image_and_batch():
    # 30 x 22: rows / columns beyond the image inside the 16 x 8 patches; depth 3 < kernel depth; two images
    _case(expect_wgrad=WG_NARROW, name="3d k5 p2 32->32 3x30x22 N=2", cin=32, cout=32, k=5, s=1, p=2, H=30, W=22, D=3, N=2)
    _case(expect_wgrad=WG_NARROW, name="3d k5 p2 16->16 7x46x38", cin=16, cout=16, k=5, s=1, p=2, H=46, W=38, D=7)


@pytest.mark.parametrize("cin,cout", [(32, 32), (16, 16), (32, 16), (16, 32)])
def test_xsplit_on_every_layer_it_supports(cin, cout):
    """knob 4 = 5: the dx taps as columns of one MMA + shift-and-add epilogue, all four channel combinations, three /
    four tiles per CTA, ragged images."""
    _case(expect_last=XSPLIT, knobs={4: 5}, name=f"3d k5 p2 {cin}->{cout} 4x64x24 (x-split)", cin=cin, cout=cout, k=5, s=1,
          p=2, H=64, W=24, D=4)
    _case(expect_last=XSPLIT, knobs={4: 5}, name=f"3d k5 p2 {cin}->{cout} 3x30x22 N=2 (x-split)", cin=cin, cout=cout, k=5,
          s=1, p=2, H=30, W=22, D=3, N=2)
    _case(expect_last=XSPLIT, knobs={4: 5}, name=f"3d k3 p1 {cin}->{cout} 5x40x28 (x-split)", cin=cin, cout=cout, k=3, s=1,
          p=1, H=40, W=28, D=5)


@pytest.mark.parametrize("cin,cout", [(32, 32), (16, 16), (32, 16), (16, 32)])
def test_one_mma_per_tap_kernel_when_xsplit_is_off(cin, cout):
    """knob 4 = 4: the same layers on igemm_halo_narrow (which otherwise only serves what x-split declines)."""
    _case(expect_last=NARROW, knobs={4: 4}, name=f"3d k5 p2 {cin}->{cout} 4x64x24 (narrow halo)", cin=cin, cout=cout, k=5,
          s=1, p=2, H=64, W=24, D=4)
    _case(expect_last=NARROW, knobs={4: 4}, name=f"3d k5 p2 {cin}->{cout} 3x30x22 (narrow halo)", cin=cin, cout=cout, k=5,
          s=1, p=2, H=30, W=22, D=3)


def test_four_patch_columns():
    """Images that columns of four 16 x 8 patches fit (the V-Net shapes): NP = 4 accumulators per CTA, single-buffered
    halo for 32 channels, ragged last column, every column-tile width."""
    _case(expect_wgrad=WG_NARROW, name="3d k5 p2 32->32 4x64x16", cin=32, cout=32, k=5, s=1, p=2, H=64, W=16, D=4)
    _case(expect_last=NARROW, knobs={4: 4}, expect_wgrad=WG_NARROW, name="3d k5 p2 16->16 3x120x24 N=2", cin=16, cout=16, k=5, s=1, p=2, H=120, W=24, D=3, N=2)
    _case(expect_wgrad=WG_NARROW, name="3d k5 p2 32->16 6x128x40", cin=32, cout=16, k=5, s=1, p=2, H=128, W=40, D=6)
    _case(expect_last=None, name="3d k5 p2 32->64 3x64x16", cin=32, cout=64, k=5, s=1, p=2, H=64, W=16, D=3)
    _case(expect_last=None, name="3d k5 p2 16->64 3x64x16", cin=16, cout=64, k=5, s=1, p=2, H=64, W=16, D=3)
    _case(expect_wgrad=WG_NARROW, name="2d k7 p3 16->32 128x32", cin=16, cout=32, k=7, s=1, p=3, H=128, W=32)
    _case(expect_last=NARROW, expect_wgrad=WG_NARROW, knobs={13: 3, 4: 4}, name="3d k5 p2 32->32 4x64x16 (one patch per CTA)",
          cin=32, cout=32, k=5, s=1, p=2, H=64, W=16, D=4)


def test_k3_and_2d_windows():
    _case(expect_wgrad=WG_NARROW, name="3d k3 p1 32->32 5x32x32", cin=32, cout=32, k=3, s=1, p=1, H=32, W=32, D=5)
    _case(expect_wgrad=WG_NARROW, name="2d k5 p2 32->32 48x40", cin=32, cout=32, k=5, s=1, p=2, H=48, W=40)
    _case(expect_last=None, expect_wgrad=WG_NARROW, name="2d k7 p3 16->16 32x32", cin=16, cout=16, k=7, s=1, p=3, H=32, W=32)


def test_column_tiles_and_partial_channels():
    # 64 output columns (BN = 64) forward; its data gradient reads 64 channels (other kernels)
    _case(expect_last=None, name="3d k5 p2 32->64 4x32x16", cin=32, cout=64, k=5, s=1, p=2, H=32, W=16, D=4)
    # 24 output channels: a partial column tile forward, padded-channel gradient (24 -> pad 24: other kernels)
    _case(expect_last=None, name="3d k5 p2 16->24 4x32x16", cin=16, cout=24, k=5, s=1, p=2, H=32, W=16, D=4)
    # 20 real channels in 32 stored ones: dOut view of 24 -> ... stays off this path; 12 -> 16 padded input on it
    _case(expect_last=None, name="3d k5 p2 12->32 4x32x16", cin=12, cout=32, k=5, s=1, p=2, H=32, W=16, D=4)


@pytest.mark.parametrize("cin,cout", [(32, 32), (16, 16), (32, 64), (16, 64)])
def test_forward_is_served_by_the_narrow_kernels(cin, cout):
    """Forward only, through ConvOp.run_fwd: path read-back + values + InstanceNorm statistics of the epilogue."""
    from ganslate_b200 import _cabi, ops
    lib = _cabi.lib()
    torch.manual_seed(1)
    op = ops.ConvOp(cin, cout, (5, 5, 5), (1, 1, 1), (2, 2, 2))
    w = (torch.randn(cout, cin, 5, 5, 5, device="cuda") * 0.05).bfloat16().float()
    b = torch.randn(cout, device="cuda") * 0.1
    x = torch.randn(2, 6, 32, 24, op.cin_pad, device="cuda").bfloat16()
    stats = torch.zeros(2, op.cout_pad, 2, device="cuda")
    lib.gb_debug_knob(15, 0)
    y = op.run_fwd(ops.make_view(x), "cuda", w, b, stats=stats)
    torch.cuda.synchronize()
    assert lib.gb_debug_knob(15, 0) == (XSPLIT if (cin == 16 and cout == 16) else NARROW)
    ref = torch.nn.functional.conv3d(x[..., :cin].float().permute(0, 4, 1, 2, 3), w, b, padding=2)
    got = y[..., :cout].float().permute(0, 4, 1, 2, 3)
    assert ((got - ref).abs().max() / ref.abs().max()).item() <= 1e-2
    # statistics of the bf16-rounded outputs, per (image, channel)
    s_ref = got.double().sum(dim=(2, 3, 4))
    q_ref = (got.double() ** 2).sum(dim=(2, 3, 4))
    assert torch.allclose(stats[:, :cout, 0].double(), s_ref, rtol=1e-3, atol=1e-2 * q_ref.sqrt().max().item())
    assert torch.allclose(stats[:, :cout, 1].double(), q_ref, rtol=1e-3)


def test_switch_off_knobs_restore_the_previous_kernels():
    _case(expect_last=1, expect_wgrad=0, knobs={4: 3, 12: 5}, name="3d k5 p2 32->32 4x32x16 (knob 4 = 3)", cin=32, cout=32, k=5, s=1, p=2, H=32, W=16, D=4)
