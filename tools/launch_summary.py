#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares of the
LAST training step in the file (the first step is warm-up). Usage: launch_summary.py launches.csv [n_steps] [--md]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            unit = row["Metric Unit"]
            v = float(row["Metric Value"].replace(",", ""))
            v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
            rows.append((row["Kernel Name"], v, row["Grid Size"]))
    return rows


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")
    name = re.sub(r"at::native::|at::", "", name)
    return name[:60]


def main():
    path = sys.argv[1]
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 2
    md = "--md" in sys.argv
    rows = load(path)
    per = len(rows) // n_steps
    step = rows[-per:]
    agg = collections.defaultdict(list)
    for name, v, g in step:
        agg[short(name)].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"launches in the step: {len(step)}, summed kernel time: {tot / 1e3:.3f} ms (serialised, profiler clocks)")
    if md:
        print("\n| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:18]:
        if md:
            print(f"| `{k}` | {len(v)} | {sum(v):.1f} | {sum(v) / len(v):.2f} | {sum(v) / tot:.3f} |")
        else:
            print(f"{k:60s} n={len(v):4d} total={sum(v):9.1f} avg={sum(v) / len(v):8.2f} share={sum(v) / tot:.3f}")
    grid = collections.defaultdict(list)
    for name, v, g in step:
        if "igemm" in name:
            grid[(short(name), g)].append(v)
    print()
    for k, v in sorted(grid.items(), key=lambda kv: -sum(kv[1]))[:14]:
        print(f"{k[0]:34s} grid {k[1]:18s} n={len(v):3d} avg={sum(v) / len(v):8.2f}us total={sum(v):8.1f}")


if __name__ == "__main__":
    main()
