"""Collates one GPU visit's artefacts (gpurun_out/<tag>/ written by tools/gpu_round.sh) into a table: every bench line
(variant -> img/s, e2e, ms/step, clocks, conv / InstanceNorm roofline fractions) and every pytest log's last line.

    python tools/summarize_round.py gpurun_out/r02a [--md]
"""
import glob
import json
import os
import sys


def last_json(path):
    try:
        for ln in reversed(open(path).read().strip().splitlines()):
            ln = ln.strip()
            if ln.startswith("{") and ln.endswith("}"):
                return json.loads(ln)
    except (OSError, ValueError):
        pass
    return None


def main():
    d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
    md = "--md" in sys.argv
    rows = []
    for path in sorted(glob.glob(os.path.join(d, "bench_*.json"))):
        line = last_json(path)
        name = os.path.basename(path)[len("bench_"):-len(".json")]
        if line is None:
            err = ""
            try:
                err = open(path[:-5] + ".err").read().strip().splitlines()[-1][:120]
            except (OSError, IndexError):
                pass
            rows.append((name, "-", "-", "-", "-", "-", "-", "no line: " + err))
            continue
        if "unavailable" in line:
            rows.append((name, "-", "-", "-", "-", "-", "-", "unavailable: " + line["unavailable"]))
            continue
        det = line.get("roofline_detail", {})
        frac = lambda k: f"{det[k]['frac']:.2f}" if k in det else "-"
        clk = line.get("clocks") or {}
        note = ",".join(clk.get("reasons", [])) if clk else ""
        if line.get("aux_errors"):
            note += " aux_errors=" + ";".join(line["aux_errors"])
        rows.append((name, f"{line['value']:.1f}", f"{line.get('e2e', {}).get('value', float('nan')):.1f}",
                     f"{line['ms_per_step']:.2f}", str(line.get("n_gpus", 1)),
                     f"{line['roofline']['frac']:.3f}" if "roofline" in line else "-",
                     f"{frac('in_fwd')}/{frac('in_bwd')}", (line.get("metric", "") + " " + note).strip()))
    hdr = ("bench", "value", "e2e", "ms/step", "gpus", "conv frac", "IN fwd/bwd frac", "metric / notes")
    if md:
        print("| " + " | ".join(hdr) + " |")
        print("|" + "---|" * len(hdr))
        for r in rows:
            print("| " + " | ".join(r) + " |")
    else:
        w = [max(len(str(x[i])) for x in rows + [hdr]) for i in range(len(hdr))]
        for r in [hdr] + rows:
            print("  ".join(str(v).ljust(w[i]) for i, v in enumerate(r)))
    print()
    for path in sorted(glob.glob(os.path.join(d, "pytest_*.log"))):
        try:
            lines = [ln for ln in open(path).read().strip().splitlines() if ln.strip()]
        except OSError:
            continue
        tail = [ln for ln in lines if " passed" in ln or " failed" in ln or "pytest exit" in ln or " error" in ln][-2:]
        print(f"{os.path.basename(path):28s} " + " | ".join(t.strip("= ") for t in tail))
    unv = sorted(glob.glob(os.path.join(os.path.dirname(d.rstrip("/")) or ".", "unverified", "*.log")))
    if unv:
        print("\nunverified tests that failed or hung (child output kept):", ", ".join(os.path.basename(u) for u in unv))


if __name__ == "__main__":
    main()
