"""Measurement script (not a test): per-tensor errors of one CycleGAN iteration on the B200 path against BOTH CPU
oracles (matched bf16-point and fp32), at several sizes.  Writes gpurun_out/<tag>/parity_matrix.json; the stated
tolerances in BASELINE.md / tests are set from these numbers.
usage: python tests/parity_matrix.py <outdir> [size:blocks ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    from parity_util import step_report
    out = sys.argv[1]
    cases = sys.argv[2:] or ["64:3", "128:9", "256:9"]
    res = {}
    for c in cases:
        size, nb = (int(v) for v in c.split(":"))
        for kind in ("matched", "fp32"):
            rep = step_report(size=size, batch=1, n_blocks=nb, step_optimizers=False, verbose=False, matched=(kind == "matched"))
            g = rep["grads"]
            w = {k: v for k, v in g.items() if k.endswith("weight")}
            res[f"{c}/{kind}"] = {
                "losses_rel": {k: abs(a - b) / max(abs(a), 1e-12) for k, (a, b) in rep["losses"].items()},
                "visuals_rel_l2": {k: v[0] for k, v in rep["visuals"].items()},
                "visuals_max_rel": {k: v[1] for k, v in rep["visuals"].items()},
                "weight_grad_rel_l2_max": max(v[0] for v in w.values()),
                "weight_grad_rel_l2_median": sorted(v[0] for v in w.values())[len(w) // 2],
                "weight_grad_max_rel_max": max(v[1] for v in w.values()),
                "weight_grad_cos_min": min(v[3] for v in w.values()),
                "weight_grads": {k: [round(x, 6) for x in v] for k, v in w.items()},
            }
            r = res[f"{c}/{kind}"]
            print(c, kind, "loss", max(r["losses_rel"].values()), "vis", r["visuals_rel_l2"], "wgrad l2 max/med",
                  r["weight_grad_rel_l2_max"], r["weight_grad_rel_l2_median"], "maxrel", r["weight_grad_max_rel_max"],
                  "cos", r["weight_grad_cos_min"], flush=True)
    os.makedirs(out, exist_ok=True)
    json.dump(res, open(os.path.join(out, "parity_matrix.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
