#!/usr/bin/env python
"""CycleGAN training throughput on B200 (BASELINE.json metric: CycleGAN train img/s at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # this framework (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

A step = one CycleGAN.optimize_parameters() (4 generator passes, G step, 2 discriminator steps, both Adam
updates, DDP gradient all-reduce when N>1) on one synthetic batch of Resnet2D-9 + PatchGAN2D at 3x256x256.
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CycleGAN train img/s"
UNIT = "img/s"

# The five configurations BASELINE.json names.  The default (config 1 at batch 8 per GPU) is the bench line the driver
# reads; the others are selected with --workload for the per-shape throughput table (`tools/gpu_round.sh ... shapes`).
#   name: (metric, preset, per-GPU batch default, input shape after the batch axis, CUDA-graph replay supported)
WORKLOADS = {
    "cyclegan2d": ("CycleGAN train img/s", "cyclegan_resnet2d", 8, None, True),
    "pix2pix_resnet": ("Pix2Pix Resnet2D train img/s", "pix2pix_resnet2d", 8, (3, 256, 512), True),
    "pix2pix_unet": ("Pix2Pix Unet2D train img/s", "pix2pix_unet2d", 8, (3, 256, 512), True),
    "cut": ("CUT train img/s", "cut_resnet2d", 1, (3, 256, 256), True),   # graph segments (r02b: 27.4 eager -> 75.9 img/s)
    "cyclegan3d": ("CycleGAN 3D Vnet3D train patches/s", "cyclegan_vnet3d", 1, (1, 32, 256, 256), True),   # r02al: 14.0 eager -> 15.8
    # (RevGAN replays two captured phases on one GPU -- r02am: 11.6 -> 12.7 and 27.3 -> 31.8 patches/s; a data-parallel
    #  run uses eager DistributedDataParallel: BaseGAN.parallelize_networks switches the capture off)
    "revgan3d": ("RevGAN 3D Vnet3D train patches/s", "revgan_vnet3d", 1, (4, 128, 128, 128), True),
    "revgan_piresnet3d": ("RevGAN 3D Piresnet3D train patches/s", "revgan_piresnet3d", 1, (1, 32, 176, 176), True),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("GB_BENCH_BATCH", 8)),
                    help="per-GPU batch (8 = the per-GPU batch BASELINE.json's GPU configs name; the reference's "
                         "own CPU case is batch 1: --batch 1)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--workload", default="cyclegan2d", choices=sorted(WORKLOADS),
                    help="which BASELINE.json configuration to time (default: the CycleGAN headline configuration)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--graph", action="store_true", help="CUDA-graph replay also for the workloads that default to eager "
                                                         "launches (CUT: capture path not yet verified on a B200)")
    ap.add_argument("--e2e-pipeline", action="store_true",
                    help="opt-in, not yet verified on a B200: e2e leg with the inputs' H2D copy on its own stream "
                         "(train.input_prefetch) and the loss of step i read after step i+1 was enqueued")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--multi-stream", action="store_true", default=None,
                    help="the two cycle chains / the two discriminators on two CUDA streams (train.multi_stream); default "
                         "for the CycleGAN workloads (measured r02k: 352 -> 376 img/s at batch 8, 108 -> 124 at batch 1)")
    ap.add_argument("--single-stream", action="store_true", help="switch train.multi_stream off")
    ap.add_argument("--no-batch1", action="store_true", help="skip the extra batch-1 leg of the headline workload")
    ap.add_argument("--roofline-all-ranks", action="store_true", help="N>1: profile the per-kernel roofline too")
    args = ap.parse_args()
    if args.workload != "cyclegan2d" and "--batch" not in sys.argv and "GB_BENCH_BATCH" not in os.environ:
        args.batch = WORKLOADS[args.workload][2]
    if args.multi_stream is None:
        args.multi_stream = (args.workload in ("cyclegan2d", "cyclegan3d", "pix2pix_resnet", "pix2pix_unet", "revgan3d", "revgan_piresnet3d", "cut")
                             and not args.single_stream)
    return args


def workload_config(args, world):
    if args.workload != "cyclegan2d":
        metric, preset, _, shape, graph = WORKLOADS[args.workload]
        return {"workload": f"{metric.replace(' train img/s', '').replace(' train patches/s', '')} "
                            f"(ganslate_b200.presets.{preset}), synthetic {'x'.join(map(str, shape))}, batch {args.batch}/GPU"
                            + (", independent passes on two CUDA streams" if args.multi_stream else ""),
                "global_batch": args.batch * world, "image": list(shape), "parallelism": f"dp{world}",
                "l2_policy": "2 x 126 MB flush buffer written between timed steps"}
    return {
        "workload": f"CycleGAN Resnet2D-9blk + PatchGAN2D(n_layers 3), synthetic 3x{args.size}x{args.size}, "
                    f"batch {args.batch}/GPU, lambda 10, lsgan, Adam(2e-4, 0.5/0.999)"
                    + (", the two cycle chains on two CUDA streams" if args.multi_stream else ""),
        "global_batch": args.batch * world, "image": [3, args.size, args.size],
        "parallelism": f"dp{world}", "l2_policy": "activations+gradients per step exceed nothing cached across "
                                                  "steps: 2 x 126 MB flush buffer written between timed steps",
    }


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_step_rate(size, batch, steps, warmup=1):
    """The reference's algorithm (oracle port, fp32, torch CPU kernels) on all host cores."""
    import torch
    from oracle import torch_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = O.OracleCycleGAN(seed=0)
    a, b = O.synthetic_batch(batch, 3, size, seed=1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        model.optimize_parameters(a, b)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return batch / med, med, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    if args.workload != "cyclegan2d":
        raise SystemExit("--impl reference times the headline configuration (--workload cyclegan2d) only")
    steps = max(1, min(args.steps, 5))
    rate, med, cores = cpu_step_rate(args.size, args.batch, steps, warmup=min(args.warmup, 1))
    world = 1
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} full CycleGAN steps (batch {args.batch}) of oracle/torch_oracle.py on the "
                                   f"host CPU (capped at 5 steps whatever --steps says); ONE {args.batch}-image CPU "
                                   f"process at every --gpus N (the CPU path has no multi-GPU form); the Python "
                                   f"reference cannot travel to the GPU box"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def conv_flops_per_step(batch):
    """12 F_G + 16 F_D - 2 dgrad1(G) - 4 dgrad1(D) at 3x256x256, per sample pair."""
    return 1287.1e9 * batch


def run_b200(args):
    import torch
    import torch.distributed as dist
    from ganslate_b200 import _cabi, profiler
    from ganslate_b200 import presets
    from ganslate_b200.utils import communication
    from ganslate_b200.utils.builders import build_gan

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        communication.init_distributed()
    dev = torch.device("cuda", local_rank)
    lib = _cabi.lib()  # fails loudly if the CUDA extension is missing

    torch.manual_seed(0)
    metric, preset, _, shape, graph_ok = WORKLOADS[args.workload]
    default_wl = args.workload == "cyclegan2d"
    if shape is None:
        shape = (3, args.size, args.size)
    if not graph_ok and not args.graph:
        args.no_graph = True  # (a workload marked eager-by-default in WORKLOADS; --graph overrides)
    flush = torch.empty(2 * 126 * 1024 * 1024, dtype=torch.uint8, device=dev)
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(batch, with_clocks):
        """Build the workload at `batch` per GPU and time args.steps resident steps, then args.steps e2e steps."""
        conf = getattr(presets, preset)(batch_size=batch, cuda_graph=not args.no_graph,
                                        **({"input_prefetch": True} if args.e2e_pipeline else {}),
                                        **({"multi_stream": True} if args.multi_stream else {}))
        model = build_gan(conf)
        # synthetic inputs U(-1, 1) (images are normalised to [-1, 1] in the reference); each rank draws its own shard
        gen = torch.Generator(device="cpu").manual_seed(1 + rank)
        a_host = torch.rand((batch,) + tuple(shape), generator=gen) * 2 - 1
        b_host = torch.rand((batch,) + tuple(shape), generator=gen) * 2 - 1
        a_host, b_host = a_host.pin_memory(), b_host.pin_memory()
        a_dev, b_dev = a_host.to(dev), b_host.to(dev)

        def step(resident=True):
            if resident:
                model.set_input({"A": a_dev, "B": b_dev})
            else:
                model.set_input({"A": a_host, "B": b_host})  # H2D from pinned memory inside the timed region
            model.optimize_parameters()

        def timed(n, resident):
            evs = []
            pending = None  # --e2e-pipeline: event after the D2H copy of the previous step's loss
            for i in range(n):
                flush.zero_()  # evict L2 between timed iterations
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step(resident)
                if not resident:
                    if args.e2e_pipeline:
                        # every step's loss still travels to the host, but the host waits for step i-1's copy only
                        # after step i has been enqueued (a tracker that logs with one step of lag)
                        loss_host[i & 1].copy_(next(iter(model.losses.values())).detach().reshape(()), non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record()
                        if pending is not None:
                            pending.synchronize()
                        pending = ev
                    else:
                        _ = float(next(iter(model.losses.values())).detach())  # D2H read of a step result
                e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            return sum(e0.elapsed_time(e1) for e0, e1 in evs) / 1e3

        # graph mode: 11 eager iterations precede the capture (PyTorch's DDP + CUDA-graph recipe), then 2 replays
        n_warm = max(args.warmup, 3) + (0 if args.no_graph else model.graph_warmup_iters + 2)
        for _ in range(n_warm):
            step()
        barrier()
        sampler = ClockSampler(local_rank) if (rank == 0 and with_clocks) else None  # rank 0's GPU is the one reported
        l0 = lib.gb_launch_count()
        barrier()
        t = timed(args.steps, True)
        barrier()
        launches = lib.gb_launch_count() - l0
        clocks = sampler.stop() if sampler is not None else None
        if getattr(model, "graph_launches_per_step", None):
            launches = model.graph_launches_per_step * args.steps
        # e2e: host buffers in, loss scalar out
        barrier()
        t_e2e = timed(args.steps, False)
        barrier()
        times = torch.tensor([t, t_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
        t, t_e2e = times.tolist()
        return dict(model=model, a_dev=a_dev, b_dev=b_dev, t=t, t_e2e=t_e2e, launches=launches, clocks=clocks,
                    n_warm=n_warm, in_bytes=2 * a_host.numel() * 4)

    if not default_wl:
        args.no_cpu_baseline = True  # the CPU leg times the headline configuration only
    m = measure(args.batch, True)
    model, a_dev, b_dev, t, t_e2e = m["model"], m["a_dev"], m["b_dev"], m["t"], m["t_e2e"]
    launches, clocks, n_warm = m["launches"], m["clocks"], m["n_warm"]

    roof = None
    aux_errors = {}
    if not args.no_roofline and (world == 1 or args.roofline_all_ranks):
        # per-kernel roofline: reported at N=1 (the kernels are the same at any N).  The profiled eager steps contain
        # the gradient all-reduce, so at N>1 EVERY rank must run them (a rank-0-only run dead-locks NCCL: r01m);
        # --roofline-all-ranks does that
        # (the throughput numbers above are already measured: a failure of an auxiliary leg must not lose the line)
        try:
            roof = profiler.conv_roofline(model, a_dev, b_dev, steps=3, with_traffic=default_wl)
        except Exception as e:  # noqa: BLE001
            if world > 1:
                raise  # the other ranks are inside the same collective sequence
            aux_errors["roofline"] = f"{type(e).__name__}: {e}"
        barrier()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            rate, med, cores = cpu_step_rate(args.size, args.batch, steps=3, warmup=1)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"3 full CycleGAN steps (batch {args.batch}) of the CPU oracle after 1 warm-up, median"}
        except Exception as e:  # noqa: BLE001
            aux_errors["cpu_baseline"] = f"{type(e).__name__}: {e}"
    batch1 = None
    if world == 1 and default_wl and args.batch != 1 and not args.no_batch1:
        # BASELINE config 1 is quoted at batch 1: the same workload at batch 1 per GPU, same harness, as extra keys
        try:
            del model
            m1 = measure(1, False)
            batch1 = {"value": args.steps / m1["t"], "unit": UNIT, "ms_per_step": m1["t"] / args.steps * 1e3,
                      "e2e": args.steps / m1["t_e2e"], "global_batch": 1}
            if args.size == 256:
                peaks = profiler.measured_peaks()
                batch1["conv_tflops_step"] = conv_flops_per_step(1) * args.steps / m1["t"] / 1e12
                batch1["frac_of_sustained_bf16_step"] = batch1["conv_tflops_step"] / peaks["bf16_tflops_sustained"]
        except Exception as e:  # noqa: BLE001
            aux_errors["batch1"] = f"{type(e).__name__}: {e}"
    if rank == 0:
        imgs = args.batch * world * args.steps
        in_bytes = m["in_bytes"]
        line = {
            "metric": metric, "value": imgs / t, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": imgs / t_e2e, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clocks,
            # (what the model actually did: a data-parallel RevGAN switches its capture off, BaseGAN.parallelize_networks)
            "cuda_graph": bool(getattr(m["model"], "use_cuda_graph", not args.no_graph)),
        }
        if args.e2e_pipeline:
            line["e2e"]["pipeline"] = "H2D of step i+1 on a copy stream; loss of step i read after step i+1 is enqueued"
        if default_wl and args.size == 256:
            line["conv_tflops_step"] = conv_flops_per_step(args.batch) * args.steps / t / 1e12
        if roof is not None:
            line["roofline"] = roof["dominant"]
            line["roofline_detail"] = roof["detail"]
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if batch1 is not None:
            line["batch1"] = batch1
        if aux_errors:
            line["aux_errors"] = aux_errors
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
