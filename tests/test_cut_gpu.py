"""CUT: PatchNCE kernel parity and one CUT iteration against the CPU oracle (fixed patch ids).

Tolerances: PatchNCE kernel (fp32 in, fp32 out) 1e-4 relative on loss and gradient; CUT losses 2e-2 relative vs the
fp32 oracle (the NCE terms depend on bf16 encoder features)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,P,D", [(1, 256, 256), (2, 64, 32), (3, 100, 256)])
def test_patchnce_kernel_matches_oracle(B, P, D):
    from ganslate_b200 import ops
    from oracle import torch_oracle as O
    torch.manual_seed(0)
    q = torch.nn.functional.normalize(torch.randn(B * P, D), dim=1)
    k = torch.nn.functional.normalize(torch.randn(B * P, D), dim=1)
    qr = q.clone().requires_grad_(True)
    ref = O.patchnce_loss(qr, k, B, 0.07)
    w = torch.rand(B * P)
    (ref * w).sum().backward()
    qg = q.cuda().requires_grad_(True)
    got = ops.PatchNCEFn.apply(qg, k.cuda(), B, 0.07)
    (got * w.cuda()).sum().backward()
    torch.cuda.synchronize()
    assert torch.allclose(got.cpu(), ref.detach(), rtol=1e-4, atol=1e-4)
    assert torch.allclose(qg.grad.cpu(), qr.grad, rtol=1e-3, atol=1e-4 * qr.grad.abs().max().item())


def test_encoder_taps_match_reference_semantics():
    """Features at encoder indices (0, 4, 8, 12, 16) incl. the padded input and the in-place ReLU effect."""
    from ganslate_b200.nn.generators import Resnet2D
    from ganslate_b200.nn.gans.unpaired.cut import extract_features
    from oracle import torch_oracle as O
    torch.manual_seed(0)
    ref = O.init_weights(O.OracleResnet2D(3, 3, 9))
    net = Resnet2D(3, 3, "instance", 9)
    net.load_state_dict(ref.state_dict())
    net = net.cuda()
    x = torch.rand(2, 3, 64, 64) * 2 - 1
    fr = O.extract_features(x, ref, (0, 4, 8, 12, 16))
    fo = extract_features(x.cuda(), net, (0, 4, 8, 12, 16))
    assert [tuple(f.shape) for f in fo] == [tuple(f.shape) for f in fr]
    assert fo[0].shape[-1] == 70 and float(fr[2].min()) == 0.0 and float(fo[2].min()) == 0.0
    for a, b, tol in zip(fo, fr, (5e-3, 1e-2, 2e-2, 3e-2, 4e-2)):
        assert ((a.cpu() - b).norm() / b.norm()).item() < tol


@pytest.mark.parametrize("multi_stream", [False, True])
def test_cut_step_vs_oracle(multi_stream):
    """(multi_stream: the translation / identity passes, the real / fake discriminator passes, the two contrastive terms
    and their two encoder passes each fork onto two CUDA streams -- same losses and gradients.)"""
    from ganslate_b200.presets import cut_resnet2d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle as O
    from parity_util import cosine
    torch.manual_seed(0)
    oracle = O.OracleCUT(n_residual_blocks=9, num_patches=256, seed=0)
    torch.manual_seed(0)
    ours = build_gan(cut_resnet2d(multi_stream=multi_stream))
    for name in ("G", "D", "mlp"):
        for (k1, p1), (k2, p2) in zip(oracle.networks[name].state_dict().items(), ours.networks[name].state_dict().items()):
            assert k1 == k2 and torch.equal(p1, p2.cpu()), (name, k1)
    a, b = O.synthetic_batch(1, 3, 128, seed=1)
    g = torch.Generator().manual_seed(5)
    sizes = [134 * 134, 64 * 64, 32 * 32, 32 * 32, 32 * 32]
    ids = [torch.randperm(s, generator=g)[:256] for s in sizes]
    lo, _ = oracle.optimize_parameters(a, b, patch_ids=ids, step_optimizers=False)
    ours.fixed_patch_ids = [i.cuda() for i in ids]
    for o in ours.optimizers.values():
        o.step = lambda *a, **k: None
    ours.set_input({"A": a, "B": b})
    ours.optimize_parameters()
    torch.cuda.synchronize()
    for k, v in lo.items():
        assert abs(float(ours.losses[k].detach()) - v) <= 2e-2 * abs(v), (k, v, float(ours.losses[k].detach()))
    # MLP gradients come through the fused PatchNCE backward; generator gradients through the feature taps
    for name in ("mlp", "G", "D"):
        po, pg = dict(oracle.networks[name].named_parameters()), dict(ours.networks[name].named_parameters())
        for k in po:
            if k.endswith("weight") and po[k].grad is not None:
                assert cosine(pg[k].grad, po[k].grad) > 0.85, (name, k, cosine(pg[k].grad, po[k].grad))


@pytest.mark.parametrize("shape,P,nc", [((1, 3, 70, 70), 256, 256), ((2, 128, 32, 32), 256, 256), ((1, 256, 16, 16), 256, 256),
                                        ((1, 24, 6, 9, 7), 100, 64), ((3, 256, 64, 64), 256, 256),
                                        ((1, 40, 12, 12), 77, 320)])
def test_patch_mlp_kernels_match_torch(shape, P, nc):
    """csrc/patch_mlp.cu (gather + Linear + ReLU + Linear + L2 norm, fp32 FMAs) against the reference's module chain
    (cut.py:262-276) evaluated by torch in fp32 on the CPU: outputs and every gradient <= 1e-4 max-relative."""
    from ganslate_b200 import ops
    g = torch.Generator().manual_seed(0)
    C = shape[1]
    feat = torch.randn(shape, generator=g)
    n_pos = feat[0, 0].numel()
    ids = torch.randperm(n_pos, generator=g)[:P]
    w1, b1 = torch.randn(nc, C, generator=g) * 0.1, torch.randn(nc, generator=g) * 0.1
    w2, b2 = torch.randn(nc, nc, generator=g) * 0.05, torch.randn(nc, generator=g) * 0.1
    dy = torch.randn(shape[0] * P, nc, generator=g)

    def ref(f, w1, b1, w2, b2):
        x = f.flatten(2).permute(0, 2, 1)[:, ids, :].flatten(0, 1)
        z = torch.relu(x @ w1.t() + b1) @ w2.t() + b2
        return z / (z.pow(2).sum(1, keepdim=True).pow(0.5) + 1e-7)

    tr = [t.clone().requires_grad_(True) for t in (feat, w1, b1, w2, b2)]
    yr = ref(*tr)
    yr.backward(dy)
    to = [t.clone().cuda().requires_grad_(True) for t in (feat, w1, b1, w2, b2)]
    yo = ops.PatchMlpFn.apply(to[0], ids.cuda(), *to[1:])
    yo.backward(dy.cuda())
    torch.cuda.synchronize()
    def mr(a, b):
        return ((a.cpu() - b).abs().max() / (b.abs().max() + 1e-30)).item()
    assert mr(yo.detach(), yr.detach()) <= 1e-4
    for name, a, b in zip(("dfeat", "dW1", "db1", "dW2", "db2"), to, tr):
        assert mr(a.grad, b.grad) <= 1e-4, (name, mr(a.grad, b.grad))


def test_cut_graph_replay_with_streams():
    """CUT with train.cuda_graph + train.multi_stream: the forked branches are captured into the two graph segments and
    replayed; losses stay finite, every network's weights move, and the replayed iteration draws new patch ids."""
    from ganslate_b200.presets import cut_resnet2d
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle as O
    torch.manual_seed(0)
    m = build_gan(cut_resnet2d(multi_stream=True, cuda_graph=True, cuda_graph_warmup=2))
    a, b = O.synthetic_batch(1, 3, 64, seed=1)
    w0 = {n: next(net.parameters()).detach().clone() for n, net in m.networks.items()}
    hist = []
    for _ in range(5):
        m.set_input({"A": a, "B": b})
        m.optimize_parameters()
        torch.cuda.synchronize()
        hist.append({k: float(v) for k, v in m.losses.items() if v is not None})
    assert len(m._graphs) == 2
    import math
    assert all(math.isfinite(v) for h in hist for v in h.values())
    for n, net in m.networks.items():
        assert (next(net.parameters()).detach() - w0[n]).abs().max().item() > 0, n
    assert hist[-1]["NCE"] != hist[-2]["NCE"]
