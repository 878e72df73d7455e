"""Shared helpers for the parity tests: build the B200 CycleGAN and the CPU oracle from the same weights."""
import random

import torch


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def max_rel(a, b):
    """max |a-b| / max |b|  (per-tensor max-relative error with the tensor's own scale as the floor)."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-20)).item()


def cosine(a, b):
    a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()


def build_pair(size=256, batch=1, n_blocks=9, seed=0, lambda_identity=0.0, matched=False):
    from oracle import torch_oracle as O
    from ganslate_b200.presets import cyclegan_resnet2d
    from ganslate_b200.utils.builders import build_gan
    random.seed(0)
    cls = O.OracleCycleGANBf16 if matched else O.OracleCycleGAN
    oracle = cls(O.default_cyclegan_conf(n_residual_blocks=n_blocks, lambda_identity=lambda_identity), seed=seed)
    torch.manual_seed(seed)
    conf = cyclegan_resnet2d(batch_size=batch, n_residual_blocks=n_blocks, lambda_identity=lambda_identity)
    ours = build_gan(conf)
    return oracle, ours


def step_report(size=256, batch=1, n_blocks=9, step_optimizers=False, lambda_identity=0.0, verbose=True,
                matched=False):
    """matched=False: fp32 oracle (what the reference computes).  matched=True: the same oracle with bf16
    rounding at the B200 path's storage points (checks the kernels, not the precision choice)."""
    from oracle import torch_oracle as O
    oracle, ours = build_pair(size, batch, n_blocks, lambda_identity=lambda_identity, matched=matched)
    rep = {}
    # same seed => same weights (init order and RNG consumption match the reference)
    worst = 0.0
    for name in oracle.networks:
        for (k1, p1), (k2, p2) in zip(oracle.networks[name].state_dict().items(),
                                      ours.networks[name].state_dict().items()):
            assert k1 == k2, (k1, k2)
            worst = max(worst, (p1 - p2.cpu()).abs().max().item())
    rep["init_max_abs_diff"] = worst
    a, b = O.synthetic_batch(batch, 3, size, seed=1)
    lo = oracle.optimize_parameters(a, b, step_optimizers=step_optimizers)
    ours.set_input({"A": a, "B": b})
    if step_optimizers:
        ours.optimize_parameters()
    else:
        for o in ours.optimizers.values():
            o.step = lambda *a, **k: None
        ours.optimize_parameters()
    torch.cuda.synchronize()
    rep["losses"] = {k: (lo[k], float(ours.losses[k].detach())) for k in lo}
    rep["visuals"] = {k: (rel_l2(ours.visuals[k], oracle.visuals[k]), max_rel(ours.visuals[k], oracle.visuals[k]))
                      for k in ("fake_B", "rec_A", "fake_A", "rec_B")}
    grads = {}
    for name in oracle.networks:
        po = dict(oracle.networks[name].named_parameters())
        pg = dict(ours.networks[name].named_parameters())
        for k in po:
            if po[k].grad is None:
                continue
            grads[f"{name}.{k}"] = (rel_l2(pg[k].grad, po[k].grad), max_rel(pg[k].grad, po[k].grad),
                                    po[k].grad.abs().max().item(), cosine(pg[k].grad, po[k].grad))
    rep["grads"] = grads
    if step_optimizers:
        w = {}
        for name in oracle.networks:
            po = dict(oracle.networks[name].named_parameters())
            pg = dict(ours.networks[name].named_parameters())
            for k in po:
                w[f"{name}.{k}"] = (pg[k].detach().cpu() - po[k].detach()).abs().max().item()
        rep["weights_max_abs_diff"] = max(w.values())
    if verbose:
        print("init diff", rep["init_max_abs_diff"])
        for k, v in rep["losses"].items():
            print(f"loss {k:8s} oracle {v[0]:.6f} ours {v[1]:.6f} rel {abs(v[0]-v[1])/max(abs(v[0]),1e-12):.2e}")
        for k, v in rep["visuals"].items():
            print(f"visual {k:7s} rel_l2 {v[0]:.3e} max_rel {v[1]:.3e}")
        gl = sorted(grads.items(), key=lambda kv: -kv[1][0])
        print("worst gradients by rel_l2 (rel_l2, max_rel, |ref|max):")
        for k, v in gl[:12]:
            print(f"  {k:42s} {v[0]:.3e} {v[1]:.3e} {v[2]:.3e}")
        print("weight gradients in module order (rel_l2, max_rel, |ref|max, cosine):")
        for k, v in grads.items():
            if k.endswith("weight") and (k.startswith("G_AB") or k.startswith("D_B")):
                print(f"  {k:42s} {v[0]:.3e} {v[1]:.3e} {v[2]:.3e} {v[3]:.5f}")
        nz = [v for k, v in grads.items() if k.endswith("weight")]
        print("median rel_l2 over non-trivial grads:", sorted(x[0] for x in nz)[len(nz) // 2], "n", len(nz))
        if step_optimizers:
            print("post-Adam weights max abs diff:", rep["weights_max_abs_diff"])
    return rep


if __name__ == "__main__":
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 9
    matched = len(sys.argv) > 3 and sys.argv[3] == "matched"
    step_report(size=size, n_blocks=nb, step_optimizers=False, matched=matched)
