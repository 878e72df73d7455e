"""Live per-kernel-family timing for bench.py's roofline object (CUDA events on the launch stream)."""
import json
import os
from collections import defaultdict

import torch

from . import ops


def measured_peaks():
    """Denominators: /root/repo/MEASURED_PEAKS.json (driver-written) or the profiling guide's fallback."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(here, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(family, batch):
    """DRAM bytes per launch of the family's dominant kernel from the committed ncu --set full capture
    (profiles/ncu_traffic_r*.json, latest round), or (None, None) when no capture of this batch size exists."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prof = os.path.join(here, "profiles")
    try:
        names = sorted(n for n in os.listdir(prof) if n.startswith("ncu_traffic_r") and n.endswith(".json"))
    except OSError:
        return None, None
    for name in reversed(names):
        with open(os.path.join(prof, name)) as f:
            d = json.load(f)
        if d.get("batch") == batch and family in d:
            e = d[family]
            return e["bytes_per_launch"], f"{e['kernel']}: {e['source']}"
    return None, None


def conv_roofline(model, a_dev, b_dev, steps=3, with_traffic=True):
    """Run `steps` EAGER training steps with every C-ABI launch bracketed by CUDA events and aggregate per
    family. Returns {"dominant": roofline object of the family with the largest share, "detail": {...}}."""
    was_graph = getattr(model, "use_cuda_graph", False)
    if was_graph:
        model.use_cuda_graph = False
    # per-launch durations are only meaningful when launches do not share the SMs with another stream's kernel: the
    # profiled steps run single-stream (train.multi_stream off), whatever the timed steps used
    train = model.conf.train
    had_ms = "multi_stream" in train
    old_ms = train.get("multi_stream", None)
    train["multi_stream"] = False
    # one un-timed eager step so packed weights / allocator are warm
    model.set_input({"A": a_dev, "B": b_dev})
    model.optimize_parameters()
    torch.cuda.synchronize()
    ops.PROFILE = []
    try:
        for _ in range(steps):
            model.set_input({"A": a_dev, "B": b_dev})
            model.optimize_parameters()
        torch.cuda.synchronize()
        records = ops.PROFILE
    finally:
        ops.PROFILE = None
        if was_graph:
            model.use_cuda_graph = True
        if had_ms:
            train["multi_stream"] = old_ms
        else:
            train._d.pop("multi_stream", None)
    agg = defaultdict(lambda: [0.0, 0.0, 0, ""])
    for fam, work, unit, e0, e1 in records:
        a = agg[fam]
        a[0] += e0.elapsed_time(e1) * 1e-3
        a[1] += work
        a[2] += 1
        a[3] = unit
    peaks = measured_peaks()
    total_t = sum(a[0] for a in agg.values())
    detail = {}
    for fam, (t, work, n, unit) in agg.items():
        if unit == "flop":
            ach, peak, u, bound = work / t / 1e12, peaks["bf16_tflops_sustained"], "TFLOP/s", "tensor"
        else:
            ach, peak, u, bound = work / t / 1e9, peaks["hbm_gbs"], "GB/s", "hbm"
        detail[fam] = {"bound": bound, "achieved": round(ach, 2), "peak": peak, "unit": u, "frac": round(ach / peak, 4),
                       "launches_per_step": n // steps, "avg_us": round(t / n * 1e6, 2),
                       "share_of_kernel_time": round(t / total_t, 4)}
    conv = [f for f in detail if f.startswith("conv")]
    t_conv = sum(agg[f][0] for f in conv)
    w_conv = sum(agg[f][1] for f in conv)
    n_conv = sum(agg[f][2] for f in conv)
    dominant = {
        "bound": "tensor", "achieved": round(w_conv / t_conv / 1e12, 2), "peak": peaks["bf16_tflops_sustained"],
        "unit": "TFLOP/s", "frac": round(w_conv / t_conv / 1e12 / peaks["bf16_tflops_sustained"], 4), "traffic": None,
        "kernel": "all convolution launches of a step: igemm_tma_kernel / igemm_pair_kernel (forward, data gradient) + "
                  "igemm_wgrad_kernel (weight gradient), plus the halo / gather kernels of the 3-channel layers",
        "peak_source": f"{peaks['source']} sustained bf16 (kernel timed inside a long step)",
        "avg_launch_us": round(t_conv / n_conv * 1e6, 2), "launches_per_step": n_conv // steps,
        "share_of_kernel_time": round(t_conv / total_t, 4),
        "how": "algorithmic FLOPs (SURVEY 8d) / CUDA-event time around each launch, eager single-stream steps",
    }
    traffic, src = ncu_traffic("conv", int(a_dev.shape[0])) if with_traffic else (None, None)
    if traffic is not None:
        dominant["traffic"], dominant["traffic_source"] = traffic, src
    return {"dominant": dominant, "detail": detail}
