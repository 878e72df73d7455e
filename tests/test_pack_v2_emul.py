"""Host run of the second-generation weight pack / weight-gradient unpack element bodies
(ganslate_b200/csrc/pack_v2_core.h, opt-in on the GPU through gb_debug_knob(28, 1)): compiled with g++
(tests/emul/pack_v2_emul.cpp) and compared BIT FOR BIT with the ABI restatement (tests/fake_cabi.py) on the pack /
unpack parameter blocks the real host code builds for strided, transposed, 3-D, pixel-window and operand-swapped
convolutions."""
import ctypes as C
import os
import subprocess

import pytest
import torch

from ganslate_b200 import _cabi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = tmp_path_factory.mktemp("emul") / "pack_v2_emul.so"
    res = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", os.path.join(HERE, "emul", "pack_v2_emul.cpp"), "-o",
                          str(out)], capture_output=True, text=True, cwd=ROOT)
    assert res.returncode == 0, res.stderr
    lib = C.CDLL(str(out))
    lib.pack_v2_emulate.argtypes = [C.POINTER(_cabi.PackParams)]
    lib.unpack_v2_emulate.argtypes = [C.POINTER(_cabi.UnpackBatch)]
    return lib


# cin, cout, kernel, stride, padding, transposed, output_padding
CONVS = [
    (256, 256, (1, 3, 3), (1, 1, 1), (0, 0, 0), False, (0, 0, 0)),     # residual block
    (64, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1), False, (0, 0, 0)),      # stride 2
    (128, 64, (1, 3, 3), (1, 2, 2), (0, 1, 1), True, (0, 1, 1)),       # transposed, 4 parity classes
    (3, 64, (1, 7, 7), (1, 1, 1), (0, 0, 0), False, (0, 0, 0)),        # pixel-window input layer
    (64, 3, (1, 7, 7), (1, 1, 1), (0, 0, 0), False, (0, 0, 0)),        # narrow output (operand-swapped wgrad)
    (16, 32, (2, 2, 2), (2, 2, 2), (0, 0, 0), False, (0, 0, 0)),       # 3-D down block
    (32, 16, (2, 2, 2), (2, 2, 2), (0, 0, 0), True, (0, 0, 0)),        # 3-D up block
    (40, 72, (3, 3, 3), (1, 1, 1), (1, 1, 1), False, (0, 0, 0)),       # ragged channel counts, 3-D
    (512, 1, (1, 4, 4), (1, 1, 1), (0, 1, 1), False, (0, 0, 0)),       # PatchGAN output
]


@pytest.mark.parametrize("conv", CONVS, ids=[f"{c[0]}to{c[1]}k{'x'.join(map(str, c[2]))}s{c[3][2]}{'T' if c[5] else ''}" for c in CONVS])
def test_pack_and_unpack_v2_equal_the_abi_restatement(emul, monkeypatch, conv):
    import fake_cabi
    fake_cabi.install(monkeypatch)
    from ganslate_b200 import ops
    monkeypatch.setattr(ops, "WINDOW_CONV", "force")
    cin, cout, k, s, p, tr, op_pad = conv
    op = ops.ConvOp(cin, cout, k, s, p, transposed=tr, output_padding=op_pad)
    T = k[0] * k[1] * k[2]
    torch.manual_seed(cin * 7 + cout)
    w = torch.randn((cin, cout, T) if tr else (cout, cin, T))
    fake = fake_cabi.FakeLib()
    for which in ("fwd", "dgrad"):
        pp, dst, n = op.pack_params(w, which)
        ref = dst.clone()
        ref.fill_(7.0)
        got = ref.clone()
        for buf, fn in ((ref, lambda q: fake.gb_pack_weights(q, None)), (got, lambda q: emul.pack_v2_emulate(C.byref(q)))):
            q = _cabi.PackParams.from_buffer_copy(bytes(pp))
            q.dst = buf.data_ptr()
            assert fn(q) == 0
        assert torch.equal(ref.view(torch.int16), got.view(torch.int16)), which
    # weight gradient: the real host path queues the unpack items; replay them through both implementations
    ext = (1 if k[0] == 1 else 5, 9, 10)
    x = torch.randn(2, ext[0], ext[1], ext[2], op.cin_pad).to(torch.bfloat16)
    if op.cin_pad != cin:
        x[..., cin:] = 0
    od, oh, ow = op.out_extent(ext)
    dy = torch.randn(2, od, oh, ow, op.cout_pad).to(torch.bfloat16)
    if op.cout_pad != cout:
        dy[..., cout:] = 0
    dyv = torch.nn.functional.pad(dy, (0, 0, ops.BWD_BORDER, ops.BWD_BORDER)) if op.bwd_window else ops.make_view(dy)
    q = ops.UnpackQueue()
    dw = op.run_wgrad(ops.make_view(x), dyv, tuple(w.shape[:2]) + tuple(k), torch.device("cpu"), pending=q)
    assert q.batch.count >= 1
    for acc in (0, 1):
        init = torch.randn_like(dw)
        outs = []
        for fn in (lambda b: fake.gb_unpack_wgrad_multi(b, None), lambda b: emul.unpack_v2_emulate(C.byref(b))):
            out = init.clone()
            b = _cabi.UnpackBatch.from_buffer_copy(bytes(q.batch))
            for i in range(b.count):
                b.item[i].dst = b.item[i].dst - dw.data_ptr() + out.data_ptr()
                b.item[i].accumulate = acc
            assert fn(b) == 0
            outs.append(out)
        assert torch.equal(outs[0], outs[1]), f"unpack accumulate={acc}"
        assert not torch.equal(outs[0], init)
