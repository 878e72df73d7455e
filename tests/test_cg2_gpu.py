"""Persistent warp-specialised convolution kernels (ganslate_b200/csrc/igemm_cg2.cu): the CTA-pair variant (tcgen05
cta_group::2, gb_debug_knob(16, 1)) and the single-CTA variant (gb_debug_knob(16, 2)).  The kernel was written after round 1's GPU budget was spent and has never run on a B200:
these tests are `experimental` (GB_EXPERIMENTAL=1 runs them).  Each case runs in its OWN process under a timeout --
a mis-signalled mbarrier shows up as a hang, and a hang must cost one case, not the suite.

    GB_EXPERIMENTAL=1 python -m pytest tests/test_cg2_gpu.py -m gpu -q

Checks per case: forward / data gradient / weight gradient against torch on the same bf16-valued inputs (tolerance
1e-2 max-relative, tests/gpu_bringup.py) and that the data calls were served by the cg2 kernel (knob 15 == 5)."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = [pytest.mark.gpu, pytest.mark.experimental]

# (kind, kwargs, expect_cg2): expect_cg2 False = the launch must fall through to the single-CTA kernels
CASES = [
    ("conv", dict(name="3x3 reflect1 256->256 64x64 N=2 (resblock)", cin=256, cout=256, k=3, s=1, p=0, H=64, W=64, N=2,
                  reflect=1), True),
    ("conv", dict(name="3x3 p1 64->64 32x32 (one tile pair + odd tail)", cin=64, cout=64, k=3, s=1, p=1, H=32, W=24, N=1), True),
    ("conv", dict(name="1x1 64->256 32x32", cin=64, cout=256, k=1, s=1, p=0, H=32, W=32), True),
    ("conv", dict(name="3x3 N=3 ragged 64->72 19x23", cin=64, cout=72, k=3, s=1, p=1, H=19, W=23, N=3), True),
    ("conv", dict(name="3x3 s2 p1 64->128 64x64", cin=64, cout=128, k=3, s=2, p=1, H=64, W=64, N=2), True),
    ("conv", dict(name="convT 3x3 s2 p1 op1 128->64 32x32", cin=128, cout=64, k=3, s=2, p=1, H=32, W=32, N=2,
                  transposed=True, op_pad=1), True),
    ("conv", dict(name="4x4 s1 p1 256->512 32x32", cin=256, cout=512, k=4, s=1, p=1, H=32, W=32), True),
    ("conv", dict(name="3d 3x3x3 p1 64->64 4x16x16", cin=64, cout=64, k=3, s=1, p=1, H=16, W=16, D=4), True),
    ("conv", dict(name="7x7 reflect3 3->64 64x64 (window)", cin=3, cout=64, k=7, s=1, p=0, H=64, W=64, reflect=3), None),
    ("resblock", {}, True),
    ("in", dict(name="IN relu C256 border1", C=256, H=16, W=16, N=2, act="relu", reflect_next=1), None),
]

DRIVER = r"""
import ctypes, json, sys, traceback
sys.path.insert(0, {here!r})
import torch
import gpu_bringup
from ganslate_b200 import _cabi
lib = _cabi.lib()
kind, kw, mode = json.loads(sys.argv[1])
lib.gb_debug_knob(16, mode)
lib.gb_debug_knob(21, 400)   # watchdog: a wait longer than 4e8 clocks (~0.2 s) reports who waited for what and traps
lib.gb_debug_knob(19, 0)
lib.gb_debug_knob(15, 0)
ROLES = ["producer waits for an empty stage", "MMA issuer waits for a drained accumulator",
         "MMA issuer waits for a full stage", "epilogue waits for a complete accumulator"]
try:
    if kind == "conv":
        ok = gpu_bringup.conv_case(**kw)
    elif kind == "resblock":
        ok = gpu_bringup.resblock_case()
    else:
        ok = gpu_bringup.in_case(kw["name"], kw["C"], kw["H"], kw["W"], kw["N"], kw["act"], kw["reflect_next"])
    torch.cuda.synchronize()
except Exception:
    traceback.print_exc()
    wd = (ctypes.c_int32 * 8)()
    if lib.gb_debug_cg2_watchdog(wd):
        print("WATCHDOG: %s; stage/accumulator %d, item %d, parity %d, CTA rank %d, block %d, thread %d"
              % (ROLES[wd[1]] if 0 <= wd[1] < 4 else wd[1], wd[2], wd[3], wd[4], wd[5], wd[6], wd[7]))
    sys.exit(3)
print("RESULT", json.dumps(dict(ok=bool(ok), last_data_path=lib.gb_debug_knob(15, 0), cg2_launches=lib.gb_debug_knob(19, 0))))
"""


@pytest.mark.parametrize("mode", [2, 1], ids=["persistent-1cta", "cta-pair"])
@pytest.mark.parametrize("kind,kw,expect", CASES, ids=[c[1].get("name", c[0]) for c in CASES])
def test_cg2_case(kind, kw, expect, mode):
    """mode 2 = persistent warp-specialised kernel on one CTA per SM (cta_group::1), mode 1 = the CTA-pair variant."""
    code = DRIVER.format(here=HERE)
    env = dict(os.environ, PYTHONPATH=os.path.dirname(HERE) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    try:
        res = subprocess.run([sys.executable, "-c", code, json.dumps([kind, kw, mode])], capture_output=True, text=True,
                             timeout=240, env=env)
    except subprocess.TimeoutExpired as e:
        pytest.fail(f"cg2 case hung (killed after 240 s): {kind} {kw}\n{(e.stdout or b'')[-2000:]}")
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-4000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("RESULT ")]
    assert line, out[-4000:]
    r = json.loads(line[-1][7:])
    assert r["ok"], out[-4000:]
    if expect is True:
        assert r["cg2_launches"] > 0, f"no launch was served by the cg2 kernel: {r}\n{out[-2000:]}"
