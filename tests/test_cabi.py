"""The C-ABI library loads on a CPU-only box and exports every symbol include/ganslate_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ganslate_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = _declared_symbols()
    for must in ("gb_conv_data", "gb_conv_wgrad", "gb_pack_weights", "gb_in_fwd", "gb_in_bwd", "gb_mse_const", "gb_l1",
                 "gb_last_error", "gb_version"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from ganslate_b200 import _build, _cabi
    _build.build()
    lib = ctypes.CDLL(str(_cabi.LIB_PATH))
    for s in _declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert set(_cabi.exported_symbols()) == set(_declared_symbols())
    assert _cabi.lib().gb_version() == 100


def test_struct_sizes_match_header():
    """ctypes mirrors must have the same size as the C structs (compiled with gcc from the header)."""
    import subprocess
    import tempfile
    from ganslate_b200 import _cabi
    src = '#include <stdio.h>\n#include "ganslate_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(gb_view),sizeof(gb_conv_class),sizeof(gb_conv_params),sizeof(gb_wgrad_params),sizeof(gb_pack_params),' \
          'sizeof(gb_in_fwd_params),sizeof(gb_in_bwd_params),sizeof(gb_unpack_item),sizeof(gb_unpack_batch),sizeof(gb_adam_item),sizeof(gb_adam_batch));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mine = [ctypes.sizeof(t) for t in (_cabi.View, _cabi.ConvClass, _cabi.ConvParams, _cabi.WgradParams, _cabi.PackParams,
                                       _cabi.InFwdParams, _cabi.InBwdParams, _cabi.UnpackItem, _cabi.UnpackBatch, _cabi.AdamItem, _cabi.AdamBatch)]
    assert sizes == mine


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing on the host."""
    import torch
    from ganslate_b200.nn.generators import Resnet2D
    net = Resnet2D(3, 3, "instance", 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 3, 16, 16))


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ganslate_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)


def test_workspace_sizes_match_what_the_host_allocates():
    """gb_workspace_bytes (host arithmetic) against the buffers ops.py allocates for gb_conv_wgrad / gb_in_bwd."""
    import ctypes as C
    from ganslate_b200 import _cabi, ops
    lib = _cabi.lib()
    n = C.c_int64(0)
    for cin, cout, k in [(256, 256, (1, 3, 3)), (3, 64, (1, 7, 7)), (64, 3, (1, 7, 7)), (16, 32, (5, 5, 5))]:
        op = ops.ConvOp(cin, cout, k, (1, 1, 1), (0, 0, 0))
        p = _cabi.WgradParams()
        p.rows, p.kpad = op.wg_rows, op.wg_kpad
        assert lib.gb_workspace_bytes(0, C.byref(p), C.byref(n)) == 0
        assert n.value == op.wg_rows_pad * op.wg_kpad * 4
    v = _cabi.View()
    v.N, v.C = 8, 256
    assert lib.gb_workspace_bytes(1, C.byref(v), C.byref(n)) == 0 and n.value == (8 * 256 * 2 + 4) * 4
    assert lib.gb_workspace_bytes(7, C.byref(v), C.byref(n)) != 0 and b"unknown operator" in lib.gb_last_error()
