"""Randomly drawn (transposed) convolutions on the GPU against torch on the same bf16-valued inputs -- the GPU twin of
tests/test_host_networks_cpu.py::test_random_convolutions_through_the_host_path (same generator, other seed).
The kernels are the verified ones; the SHAPES are new (odd channel counts, strides up to 3, anisotropic 3-D kernels,
output padding): `experimental` until its first GPU visit (GB_EXPERIMENTAL=1; tools/gpu_round.sh ... exp runs it)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_host_networks_cpu import _random_conv_cases  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.experimental]


@pytest.mark.parametrize("case", _random_conv_cases(24, 31337), ids=lambda c: "{}{}d k{} s{} p{} {}->{}".format(
    "T" if c["transposed"] else "", c["dims"], "x".join(map(str, c["k"])), "x".join(map(str, c["s"])),
    "x".join(map(str, c["p"])), c["cin"], c["cout"]))
def test_random_convolutions_on_the_gpu(case):
    from ganslate_b200.nn import layers
    from parity_util import max_rel
    torch.manual_seed(7)
    c = case
    F = torch.nn.functional
    if c["transposed"]:
        cls = layers.ConvTranspose3d if c["dims"] == 3 else layers.ConvTranspose2d
        conv = cls(c["cin"], c["cout"], c["k"], stride=c["s"], padding=c["p"], output_padding=c["op"], bias=c["bias"])
        fn = F.conv_transpose3d if c["dims"] == 3 else F.conv_transpose2d
        kw = dict(stride=c["s"], padding=c["p"], output_padding=c["op"])
    else:
        cls = layers.Conv3d if c["dims"] == 3 else layers.Conv2d
        conv = cls(c["cin"], c["cout"], c["k"], stride=c["s"], padding=c["p"], bias=c["bias"])
        fn = F.conv3d if c["dims"] == 3 else F.conv2d
        kw = dict(stride=c["s"], padding=c["p"])
    conv = conv.cuda()
    with torch.no_grad():
        conv.weight.copy_((torch.randn_like(conv.weight) * 0.1).to(torch.bfloat16).float())
        if c["bias"]:
            conv.bias.copy_(torch.randn_like(conv.bias) * 0.1)
    nclass = 1
    for ss in c["s"]:
        nclass *= ss
    if nclass > 8:
        with pytest.raises(ValueError, match="parity classes"):
            conv.conv_op()
        return

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = torch.nn.Sequential(conv)

        def forward(self, t):
            return layers.run_network(self, list(self.model), t)

    x = torch.randn((c["N"], c["cin"]) + c["ext"], device="cuda").to(torch.bfloat16).float().requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    wr = conv.weight.detach().clone().requires_grad_(True)
    br = conv.bias.detach().clone().requires_grad_(True) if c["bias"] else None
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        yr = fn(xr, wr, br, **kw)
        if min(yr.shape) == 0:
            pytest.skip("empty output")
        y = Net()(x)
        assert y.shape == yr.shape
        g = torch.randn_like(yr).to(torch.bfloat16).float()
        y.backward(g)
        yr.backward(g)
    torch.cuda.synchronize()
    assert max_rel(y, yr) < 1e-2, max_rel(y, yr)
    assert max_rel(x.grad, xr.grad) < 1e-2, max_rel(x.grad, xr.grad)
    assert max_rel(conv.weight.grad, wr.grad) < 1e-2, max_rel(conv.weight.grad, wr.grad)
    if c["bias"]:
        assert max_rel(conv.bias.grad, br.grad) < 1e-2
