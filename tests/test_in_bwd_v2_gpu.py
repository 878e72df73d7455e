"""Second-generation (ganslate_b200/csrc/instnorm_v2.cu, gb_debug_knob(22, 1 | 2)) and on-chip cluster
(instnorm_v3.cu, gb_debug_knob(24, 1 | 2)) fused InstanceNorm backward on the GPU.  Written after round 1's GPU budget was spent: `experimental` (GB_EXPERIMENTAL=1 runs it; tools/gpu_round.sh
<tag> in2).  Its per-thread body already runs on the CPU (tests/test_in_bwd_v2_emul.py); what this adds is the block
reduction, the atomics and the grid barrier of the single-launch form (v2), the warp / block reduction and the
distributed-shared-memory exchange of the cluster (v3) -- a barrier that is mis-counted spins forever, so every
variant runs in its own process under a timeout.

Each case calls gb_in_bwd three times on the same inputs: general kernel (knob 7 = 1), first-generation fast kernel
(default) and the variant under test, and compares dx / bias gradient / residual gradient (tolerance: bf16 rounding
of values of the gradient's scale -- the algebra is associated differently) and checks the variant was really served
by the new kernel (knob 23 counts its launches)."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = [pytest.mark.gpu, pytest.mark.experimental]

DRIVER = r"""
import ctypes as C, json, sys
sys.path.insert(0, {here!r})
import torch
from ganslate_b200 import _cabi, ops
from ganslate_b200._cabi import ACT_LEAKY, ACT_NONE, ACT_PRELU, ACT_RELU
lib = _cabi.lib()
variant_knobs = {{int(k): v for k, v in json.loads(sys.argv[1]).items()}}
counter = 25 if 24 in variant_knobs else 23
dev = "cuda"
# N, D, H, W, C, gradient border, act, residual gradient
CASES = [(2, 1, 12, 10, 64, 1, ACT_RELU, False), (1, 1, 9, 7, 8, 0, ACT_LEAKY, False), (3, 1, 8, 8, 256, 1, ACT_NONE, True),
         (2, 3, 5, 6, 16, 0, ACT_RELU, False), (1, 1, 11, 13, 24, 1, ACT_LEAKY, False), (2, 1, 8, 9, 32, 2, ACT_RELU, True),
         (8, 1, 64, 64, 256, 1, ACT_RELU, False), (8, 1, 64, 64, 256, 1, ACT_NONE, True), (4, 1, 128, 128, 64, 3, ACT_RELU, False),
         (2, 1, 31, 31, 512, 0, ACT_LEAKY, False), (600, 1, 7, 7, 16, 1, ACT_RELU, False),
         (1, 1, 90, 91, 32, 1, ACT_LEAKY, False), (8, 1, 32, 32, 256, 0, ACT_LEAKY, False), (2, 1, 64, 64, 128, 0, ACT_LEAKY, True)]
# the general form of the second generation (V-Net layers): (prelu gradient, residual before the activation, out_scale)
GEN = {{}}
lean = variant_knobs.get(22) == 4   # third generation (instnorm_fast.cu, in_bwd_lean_kernel): plain forms only
if counter == 23 and not lean:
    for case, gen in [((2, 8, 16, 16, 16, 0, ACT_PRELU, False), (True, False, 0.0)), ((1, 4, 32, 32, 32, 0, ACT_PRELU, True), (True, True, 0.0)),
                      ((2, 1, 24, 20, 64, 1, ACT_RELU, True), (False, True, 0.0)), ((1, 6, 12, 12, 16, 0, ACT_PRELU, True), (True, False, -1.0)),
                      ((1, 16, 64, 64, 32, 0, ACT_PRELU, True), (True, True, 0.0))]:
        CASES.append(case)
        GEN[case] = gen
bad = 0
for case in CASES:
    (N, D, H, W, Cc, gp, act, res) = case
    want_dprelu, rba, oscale = GEN.get(case, (False, False, 0.0))
    torch.manual_seed(N * 1000 + Cc)
    x = (torch.randn(N, D, H, W, Cc, device=dev) * 1.5 + 0.7).to(torch.bfloat16)
    dy = torch.randn(N, D, H + 2 * gp, W + 2 * gp, Cc, device=dev)
    xf = x.float()
    stats = torch.stack([xf.sum(dim=(1, 2, 3)), (xf * xf).sum(dim=(1, 2, 3))], dim=-1).contiguous()
    sum0 = torch.randn(N, D, H, W, Cc, device=dev)
    prelu = (torch.rand(Cc, device=dev) * 0.5 + 0.05) if act == ACT_PRELU else None
    res_t = (torch.randn(N, D, H, W, Cc, device=dev) * 0.8).to(torch.bfloat16) if rba else None
    outs = []
    for knobs in ({{7: 1, 22: 0, 6: 0, 24: 0}}, {{7: 0, 22: 0, 6: 2, 24: 0}}, {{**{{7: 0, 22: 0, 6: 0, 24: 0}}, **variant_knobs}}):
        for k, v in knobs.items():
            lib.gb_debug_knob(k, v)
        lib.gb_debug_knob(counter, 0)
        dx = torch.full_like(x, float("nan"))
        dsum = sum0.clone()
        bstats = torch.zeros(N * Cc * 2 + 4, device=dev)
        dbias = torch.zeros(Cc, device=dev)
        p = _cabi.InBwdParams()
        p.x, p.dy_b, p.dx = ops.make_view(x), ops.make_view(dy, gp), ops.make_view(dx)
        if res:
            p.dy_sum, p.dy_sum_acc = ops.make_view(dsum), 1
        p.stats, p.bstats, p.dbias = stats.data_ptr(), bstats.data_ptr(), dbias.data_ptr()
        p.eps, p.act, p.act_slope = 1e-5, act, 0.2 if act == ACT_LEAKY else 0.0
        dprelu = torch.zeros(Cc, device=dev)
        if prelu is not None:
            p.prelu = prelu.data_ptr()
            if want_dprelu:
                p.dprelu = dprelu.data_ptr()
        if res_t is not None:
            p.res, p.res_before_act = ops.make_view(res_t), 1
        p.out_scale = oscale
        _cabi.check(lib.gb_in_bwd(C.byref(p), torch.cuda.current_stream().cuda_stream), "gb_in_bwd")
        torch.cuda.synchronize()
        outs.append((dx.float(), dbias, dsum, lib.gb_debug_knob(counter, 0), dprelu))
    ref, gen1, gen2 = outs
    scale = ref[0].abs().max().item()
    e1 = (gen1[0] - ref[0]).abs().max().item() / scale
    e2 = (gen2[0] - ref[0]).abs().max().item() / scale
    ok = (not torch.isnan(gen2[0]).any().item()) and e2 <= 2.0 ** -7
    ok = ok and torch.allclose(gen2[1], ref[1], rtol=2e-3, atol=2e-3 * max(1.0, ref[1].abs().max().item()))
    ok = ok and (not res or torch.allclose(gen2[2], ref[2], rtol=1e-5, atol=1e-5))
    eligible = counter == 23 or (Cc % 32 == 0 and D * H * W <= 8192)   # the on-chip kernel declines the rest
    if lean:
        eligible = Cc % 8 == 0 and 256 % (Cc // 8) == 0
    ok = ok and gen2[3] == (1 if eligible else 0) and gen1[3] == 0
    ok = ok and torch.allclose(gen2[4], ref[4], rtol=2e-3, atol=2e-3 * max(1.0, ref[4].abs().max().item()))
    print(("OK  " if ok else "FAIL"), (N, D, H, W, Cc, gp, act, res), "rel dx err gen1 %.2e gen2 %.2e served %d" % (e1, e2, gen2[3]))
    bad += 0 if ok else 1
print("RESULT", json.dumps(dict(bad=bad)))
sys.exit(1 if bad else 0)
"""


@pytest.mark.parametrize("knobs", [{22: 4, 6: 2}, {22: 4, 6: 1}, {22: 1, 6: 2}, {22: 1, 6: 1}, {22: 2, 6: 2}, {22: 2, 6: 1}, {24: 1}, {24: 2}],
                         ids=["lean-fused", "lean-two-launch", "v2-U4-fused", "v2-U4-two-launch", "v2-U2-fused",
                              "v2-U2-two-launch", "v3-onchip", "v3-onchip-half-stash"])
def test_in_bwd_variant(knobs):
    code = DRIVER.format(here=HERE)
    env = dict(os.environ, PYTHONPATH=os.path.dirname(HERE) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    try:
        res = subprocess.run([sys.executable, "-c", code, json.dumps(knobs)], capture_output=True,
                             text=True, timeout=300, env=env)
    except subprocess.TimeoutExpired as e:
        out = e.stdout.decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")
        pytest.fail(f"hung (killed after 300 s), output so far:\n{out[-3000:]}")
    print(res.stdout[-4000:])
    assert res.returncode == 0, res.stdout[-4000:] + "\n" + res.stderr[-3000:]


FWD_DRIVER = r"""
import ctypes as C, json, sys
sys.path.insert(0, {here!r})
import torch
from ganslate_b200 import _cabi, ops
from ganslate_b200._cabi import ACT_LEAKY, ACT_NONE, ACT_PRELU, ACT_RELU
lib = _cabi.lib()
dev = "cuda"
# N, D, H, W, C, border of y, act, residual, (residual before the activation, out_scale, no normalisation)
CASES = [(2, 1, 12, 10, 64, 1, ACT_RELU, False, (False, 0.0, False)), (3, 1, 8, 8, 256, 1, ACT_NONE, True, (False, 0.0, False)),
         (1, 1, 11, 13, 24, 3, ACT_LEAKY, False, (False, 0.0, False)), (2, 3, 5, 6, 16, 0, ACT_RELU, False, (False, 0.0, False)),
         (1, 1, 9, 7, 8, 0, ACT_RELU, False, (False, 0.0, True)), (8, 1, 64, 64, 256, 1, ACT_RELU, False, (False, 0.0, False)),
         (8, 1, 64, 64, 256, 1, ACT_NONE, True, (False, 0.0, False)), (4, 1, 128, 128, 64, 3, ACT_RELU, False, (False, 0.0, False)),
         (2, 2, 6, 7, 16, 0, ACT_PRELU, False, (False, 0.0, False)), (1, 2, 9, 8, 32, 0, ACT_PRELU, True, (True, 0.0, False)),
         (1, 4, 5, 5, 16, 0, ACT_PRELU, True, (False, -1.0, False)), (1, 16, 64, 64, 32, 0, ACT_PRELU, True, (True, 0.0, False)),
         (600, 1, 7, 7, 16, 1, ACT_RELU, False, (False, 0.0, False))]
bad = 0
for (N, D, H, W, Cc, yp, act, res, (rba, oscale, no_norm)) in CASES:
    torch.manual_seed(N * 1000 + Cc)
    x = (torch.randn(N, D, H, W, Cc, device=dev) * 1.5 + 0.7).to(torch.bfloat16)
    xf = x.float()
    stats = torch.stack([xf.sum(dim=(1, 2, 3)), (xf * xf).sum(dim=(1, 2, 3))], dim=-1).contiguous()
    prelu = (torch.rand(Cc, device=dev) * 0.5 + 0.05) if act == ACT_PRELU else None
    res_t = (torch.randn(N, D, H, W, Cc, device=dev) * 0.8).to(torch.bfloat16) if res else None
    outs = []
    for knobs in ({{7: 1, 26: 0}}, {{7: 0, 26: 1}}):
        for k, v in knobs.items():
            lib.gb_debug_knob(k, v)
        lib.gb_debug_knob(27, 0)
        y = torch.full((N, D, H + 2 * yp, W + 2 * yp, Cc), float("nan"), device=dev).to(torch.bfloat16)
        p = _cabi.InFwdParams()
        p.x, p.y = ops.make_view(x), ops.make_view(y, yp)
        if res_t is not None:
            p.res, p.res_before_act = ops.make_view(res_t), 1 if rba else 0
        p.stats = None if no_norm else stats.data_ptr()
        p.prelu = prelu.data_ptr() if prelu is not None else None
        p.eps, p.act, p.act_slope, p.out_scale = 1e-5, act, 0.2 if act == ACT_LEAKY else 0.0, oscale
        _cabi.check(lib.gb_in_fwd(C.byref(p), torch.cuda.current_stream().cuda_stream), "gb_in_fwd")
        torch.cuda.synchronize()
        outs.append((y.float(), lib.gb_debug_knob(27, 0)))
    (ref, _), (got, served) = outs
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item() / scale
    ok = (not torch.isnan(got).any().item()) and err <= 2.0 ** -7 and served == 1
    print(("OK  " if ok else "FAIL"), (N, D, H, W, Cc, yp, act, res, rba, oscale, no_norm), "rel err %.2e served %d" % (err, served))
    bad += 0 if ok else 1
print("RESULT", json.dumps(dict(bad=bad)))
sys.exit(1 if bad else 0)
"""


def test_in_fwd_v2():
    """Second-generation forward (gb_debug_knob(26, 1)): plain and V-Net forms against the general kernel, incl. the
    reflection border of the result."""
    code = FWD_DRIVER.format(here=HERE)
    env = dict(os.environ, PYTHONPATH=os.path.dirname(HERE) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    try:
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    except subprocess.TimeoutExpired as e:
        out = e.stdout.decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")
        pytest.fail(f"hung (killed after 300 s), output so far:\n{out[-3000:]}")
    print(res.stdout[-4000:])
    assert res.returncode == 0, res.stdout[-4000:] + "\n" + res.stderr[-3000:]
