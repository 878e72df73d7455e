#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) as markdown: duration, DRAM traffic, tensor-pipe and memory
throughput, occupancy and the top warp-stall reasons per captured kernel. Usage: ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active (% of peak)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput (% of peak)"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (% of peak)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (%)"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__waves_per_multiprocessor", "waves per SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"### `{r[idx['Kernel Name']][:90]}` grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
        print("| metric | value |\n|---|---|")
        for key, label in WANT:
            if key in idx:
                print(f"| {label} | {r[idx[key]]} {units[idx[key]]} |")
        stalls = []
        for h, i in idx.items():
            if "pcsamp_warps_issue_stalled_" in h and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h.split("stalled_")[1]))
                except ValueError:
                    pass
        tot = sum(s for s, _ in stalls) or 1.0
        top = ", ".join(f"{n} {100 * s / tot:.0f}%" for s, n in sorted(stalls, reverse=True)[:5])
        print(f"| top warp-stall samples | {top} |\n")


if __name__ == "__main__":
    main()
