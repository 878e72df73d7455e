"""The persistent TMA-fed convolution kernel (igemm_pers_kernel in ganslate_b200/csrc/igemm_tma.cu, gb_debug_knob(16, 3)):
forward / data gradient / weight gradient of every layer family against torch on the same bf16-valued inputs
(tolerance 1e-2 max-relative, tests/gpu_bringup.py), each case in its OWN process under a timeout (a mis-signalled
mbarrier is a hang, and a hang must cost one case, not the suite), and a check that the data calls really were served
by the persistent kernel (knob 15 == 7)."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu

CASES = [
    ("conv", dict(name="3x3 reflect1 256->256 64x64 N=2 (resblock)", cin=256, cout=256, k=3, s=1, p=0, H=64, W=64, N=2, reflect=1)),
    ("conv", dict(name="3x3 reflect1 256->256 64x64 N=8", cin=256, cout=256, k=3, s=1, p=0, H=64, W=64, N=8, reflect=1)),
    ("conv", dict(name="3x3 p1 64->64 32x24", cin=64, cout=64, k=3, s=1, p=1, H=32, W=24, N=1)),
    ("conv", dict(name="1x1 64->256 32x32", cin=64, cout=256, k=1, s=1, p=0, H=32, W=32)),
    ("conv", dict(name="3x3 N=3 ragged 64->128 19x23", cin=64, cout=128, k=3, s=1, p=1, H=19, W=23, N=3)),
    ("conv", dict(name="3x3 s2 p1 64->128 64x64", cin=64, cout=128, k=3, s=2, p=1, H=64, W=64, N=2)),
    ("conv", dict(name="3x3 s2 p1 128->256 128x128 N=4", cin=128, cout=256, k=3, s=2, p=1, H=128, W=128, N=4)),
    ("conv", dict(name="convT 3x3 s2 p1 op1 128->64 32x32", cin=128, cout=64, k=3, s=2, p=1, H=32, W=32, N=2, transposed=True, op_pad=1)),
    ("conv", dict(name="convT 3x3 s2 p1 op1 256->128 64x64 N=4", cin=256, cout=128, k=3, s=2, p=1, H=64, W=64, N=4, transposed=True, op_pad=1)),
    ("conv", dict(name="4x4 s2 p1 64->128 64x64 (PatchGAN)", cin=64, cout=128, k=4, s=2, p=1, H=64, W=64, N=2)),
    ("conv", dict(name="4x4 s1 p1 256->512 32x32", cin=256, cout=512, k=4, s=1, p=1, H=32, W=32)),
    ("conv", dict(name="4x4 s1 p1 512->64 31x31 N=2", cin=512, cout=64, k=4, s=1, p=1, H=31, W=31, N=2)),
    ("conv", dict(name="3d 3x3x3 p1 64->64 4x16x16", cin=64, cout=64, k=3, s=1, p=1, H=16, W=16, D=4)),
    ("conv", dict(name="3d 4x4x4 s2 p1 64->128 8x32x32", cin=64, cout=128, k=4, s=2, p=1, H=32, W=32, D=8)),
    ("conv", dict(name="7x7 reflect3 3->64 64x64 (window)", cin=3, cout=64, k=7, s=1, p=0, H=64, W=64, reflect=3)),
    ("conv", dict(name="3x3 p1 64->64 tanh epilogue", cin=64, cout=64, k=3, s=1, p=1, H=20, W=20, act="tanh")),
    ("resblock", {}),
]

DRIVER = r"""
import json, sys, traceback
sys.path.insert(0, {here!r})
import torch
import gpu_bringup
from ganslate_b200 import _cabi
lib = _cabi.lib()
kind, kw = json.loads(sys.argv[1])
lib.gb_debug_knob(16, 3)
lib.gb_debug_knob(9, 1)
seen = []
orig = lib.gb_conv_data
def spy(*a):
    lib.gb_debug_knob(15, 0)
    r = orig(*a)
    seen.append(lib.gb_debug_knob(15, 0))
    return r
lib.gb_conv_data = spy
try:
    ok = gpu_bringup.conv_case(**kw) if kind == "conv" else gpu_bringup.resblock_case()
    torch.cuda.synchronize()
except Exception:
    traceback.print_exc()
    sys.exit(3)
print("RESULT", json.dumps(dict(ok=bool(ok), paths=seen)))
"""


@pytest.mark.parametrize("kind,kw", CASES, ids=[c[1].get("name", c[0]) for c in CASES])
def test_persistent_kernel_case(kind, kw):
    code = DRIVER.format(here=HERE)
    env = dict(os.environ, PYTHONPATH=os.path.dirname(HERE) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    try:
        res = subprocess.run([sys.executable, "-c", code, json.dumps([kind, kw])], capture_output=True, text=True,
                             timeout=150, env=env)
    except subprocess.TimeoutExpired as e:
        pytest.fail(f"persistent-kernel case hung (killed after 150 s): {kind} {kw}\n{(e.stdout or b'')[-2000:]}")
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-4000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("RESULT ")]
    assert line, out[-4000:]
    r = json.loads(line[-1][7:])
    assert r["ok"], out[-4000:]
    assert 7 in r["paths"], f"no data launch was served by the persistent kernel: {r}\n{out[-2000:]}"
