"""FusedAdam (csrc/adam.cu) against torch.optim.Adam on the same parameters and gradients: the update must agree
to fp32 round-off (tolerance 2e-6 relative to the parameter scale per step), and the state_dict must round-trip
through torch.optim.Adam (the checkpoint format of ganslate/nn/gans/base.py:244-245)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch_adam():
    from ganslate_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(64, 3, 7, 7), (64,), (256, 256, 3, 3), (1, 512, 4, 4), (5,), (3, 64, 7, 7)]
    ref_p = [torch.randn(s, device="cuda").mul_(0.02).requires_grad_(True) for s in shapes]
    our_p = [p.detach().clone().requires_grad_(True) for p in ref_p]
    ref = torch.optim.Adam(ref_p, lr=2e-4, betas=(0.5, 0.999))
    ours = FusedAdam(our_p, lr=2e-4, betas=(0.5, 0.999))
    for it in range(4):
        for a, b in zip(ref_p, our_p):
            g = torch.randn_like(a) * (10.0 ** (it - 2))
            a.grad, b.grad = g.clone(), g.clone()
        if it == 2:
            for grp in list(ref.param_groups) + list(ours.param_groups):
                grp["lr"] = 1e-4   # what LambdaLR does between iterations
        ref.step()
        versions = [b._version for b in our_p]
        ours.step()
        torch.cuda.synchronize()
        # raw-pointer writes must still bump the version counters (the packed-weight cache keys on them)
        assert all(b._version > v for b, v in zip(our_p, versions))
        for a, b in zip(ref_p, our_p):
            assert (a - b).abs().max().item() <= 2e-6 * max(a.abs().max().item(), 1e-3), it
    sd = ours.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 4.0
    fresh = torch.optim.Adam([p.detach().clone().requires_grad_(True) for p in our_p], lr=2e-4, betas=(0.5, 0.999))
    fresh.load_state_dict(sd)   # torch's Adam accepts the checkpoint
    again = FusedAdam(our_p, lr=2e-4, betas=(0.5, 0.999))
    again.load_state_dict(ref.state_dict())  # and FusedAdam accepts torch's
    for b in our_p:
        b.grad = torch.ones_like(b)
    again.step()
    torch.cuda.synchronize()
    assert float(again.state[our_p[0]]["step"]) == 5.0
