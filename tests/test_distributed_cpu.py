"""world_size-2 gloo tests of the N>1 host path: process-group helpers, DDP gradient averaging as wired by
BaseGAN.parallelize_networks (one DDP wrapper per network, broadcast_buffers=False), and the flat-bucket gradient
all-reduce that replaces DDP's hooks when the step is replayed as CUDA graphs (utils/grad_sync.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from ganslate_b200.utils import communication as comm
    from oracle import torch_oracle as O
    comm.init_distributed()  # gloo on a CPU box
    assert comm.get_world_size() == world and comm.get_rank() == rank and comm.get_local_rank() == rank
    seed = comm.shared_random_seed()
    red = comm.reduce({"a": torch.tensor(float(rank + 1)), "b": torch.tensor(2.0)}, average=True, all_reduce=True)
    # DDP over a small discriminator: averaged gradient == mean of the per-rank gradients
    torch.manual_seed(0)
    net = O.init_weights(O.OraclePatchGAN2D(3, 8, 2))
    ddp = torch.nn.parallel.DistributedDataParallel(net, broadcast_buffers=False)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand((1, 3, 32, 32), generator=g)
    O.adversarial_lsgan(ddp(x), True).backward()
    grad = net.model[0].weight.grad.clone()
    # reference value: both shards on one process
    torch.manual_seed(0)
    ref = O.init_weights(O.OraclePatchGAN2D(3, 8, 2))
    tot = 0
    for r in range(world):
        gr = torch.Generator().manual_seed(100 + r)
        tot = tot + O.adversarial_lsgan(ref(torch.rand((1, 3, 32, 32), generator=gr)), True)
    (tot / world).backward()
    # flat-bucket sync (the CUDA-graph path's replacement of DDP): rank-0 parameters at start, gradients averaged over
    # ranks, a parameter shared by two groups counted once, a parameter without gradient reduced as zero
    from ganslate_b200.utils.grad_sync import FlatGradSync
    torch.manual_seed(10 + rank)  # different initial weights per rank: broadcast must make them rank 0's
    net2 = O.init_weights(O.OraclePatchGAN2D(3, 8, 2))
    params = list(net2.parameters())
    sync = FlatGradSync(params + params[:2], torch.device("cpu"))
    sync.broadcast_parameters()
    torch.manual_seed(10)
    ref2 = O.init_weights(O.OraclePatchGAN2D(3, 8, 2))
    bcast_err = max(float((a - b).abs().max()) for a, b in zip(net2.parameters(), ref2.parameters()))
    O.adversarial_lsgan(net2(x), True).backward()
    params[-1].grad = None  # e.g. a frozen / unused parameter on this rank
    sync.finish()           # nothing pending: must be a no-op
    sync.launch()
    sync.finish()
    tot = 0
    for r in range(world):
        gr = torch.Generator().manual_seed(100 + r)
        tot = tot + O.adversarial_lsgan(ref2(torch.rand((1, 3, 32, 32), generator=gr)), True)
    (tot / world).backward()
    flat_err = max(float((a.grad - b.grad).abs().max()) for a, b in list(zip(net2.parameters(), ref2.parameters()))[:-1])
    ok_none = params[-1].grad is None and len(sync.params) == len(params)
    # bf16 wire format (GB_SYNC_BF16): same average to bf16 accuracy
    g32 = [p.grad.clone() for p in params[:-1]]
    net3 = O.init_weights(O.OraclePatchGAN2D(3, 8, 2))
    net3.load_state_dict(ref2.state_dict())
    O.adversarial_lsgan(net3(x), True).backward()
    sync16 = FlatGradSync(list(net3.parameters()), torch.device("cpu"), dtype=torch.bfloat16)
    sync16.launch()
    sync16.finish()
    bf16_err = max(float((a.grad - b).abs().max()) / max(1e-12, float(b.abs().max())) for a, b in zip(list(net3.parameters())[:-1], g32))
    ok_none = ok_none and bf16_err < 2.0 ** -6 and all(p.grad.dtype == torch.float32 for p in net3.parameters())
    q.put((rank, seed, float(red["a"]), float(red["b"]), float((grad - ref.model[0].weight.grad).abs().max()), bcast_err,
           flat_err, ok_none))
    comm.synchronize()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gloo_world2_helpers_and_ddp_average():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]                      # shared seed agreed by broadcast
    assert res[0][2] == pytest.approx(1.5) and res[0][3] == pytest.approx(2.0)
    assert all(r[4] < 1e-6 for r in res)               # DDP average == single-process mean
    assert all(r[5] == 0.0 for r in res)               # FlatGradSync: parameters are rank 0's after the broadcast
    assert all(r[6] < 1e-6 and r[7] for r in res)      # flat-bucket average == single-process mean


def _graph_path_worker(rank, world, port, q, recipe="cyclegan"):
    """The data-parallel CUDA-graph code path of CycleGAN.optimize_parameters (explicit flat-bucket all-reduces between
    the captured segments) on two gloo ranks through the pointer-level CPU restatement of the ABI; the capture itself
    is replaced by a direct call.  Each rank trains on its own batch; the gradients it ends up with must be the mean of
    the two single-rank gradients, and the optimizers must step in the reference's order."""
    import contextlib
    import random
    import sys
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import fake_cabi
    from ganslate_b200 import _cabi, ops
    from ganslate_b200.nn.gans import base
    from ganslate_b200.presets import cut_resnet2d, cyclegan_resnet2d, pix2pix_resnet2d
    from ganslate_b200.utils import communication as comm
    from ganslate_b200.utils.builders import build_gan
    from oracle import torch_oracle as O

    class MP:  # monkeypatch stand-in (the process ends with the test)
        def setattr(self, o, n, v, raising=True):
            setattr(o, n, v)

    fake_cabi.install(MP())
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    base.BaseGAN._specify_device = lambda self: torch.device("cpu")
    base.BaseGAN.eager_stream = lambda self: contextlib.nullcontext()
    comm.init_distributed()

    # eager data parallelism wraps every network in DistributedDataParallel(device_ids=[device]) as the reference does
    # (base.py:180-183); CPU modules take no device_ids
    real_ddp = base.DistributedDataParallel
    base.DistributedDataParallel = lambda net, device_ids=None, output_device=None, broadcast_buffers=False: real_ddp(
        net, broadcast_buffers=broadcast_buffers)

    def run(distributed, batch_seed, mode):
        random.seed(0)
        torch.manual_seed(0)
        if not distributed:  # single-rank reference: no gradient sync objects are created
            saved = dist.is_initialized
            torch.distributed.is_initialized = lambda: False
        try:
            kw = dict(batch_size=1, n_residual_blocks=1, cuda_graph=mode != "ddp", cuda_graph_warmup=0 if mode == "segments" else 100)
            if recipe == "cut":  # (its feature taps need the 9-block encoder)
                kw["n_residual_blocks"] = 9
                conf = cut_resnet2d(**kw)
                conf.train.gan.optimizer.num_patches = 16
                gan = build_gan(conf)
                g5 = torch.Generator().manual_seed(5)
                gan.fixed_patch_ids = [torch.randperm(n, generator=g5)[:16] for n in [54 * 54, 24 * 24, 12 * 12, 12 * 12, 12 * 12]]
            else:
                gan = build_gan(cyclegan_resnet2d(**kw) if recipe == "cyclegan" else pix2pix_resnet2d(n_layers=3, **kw))
        finally:
            if not distributed:
                torch.distributed.is_initialized = saved
        order = []
        for name, o in gan.optimizers.items():
            o.step = lambda *a, _n=name, **k: order.append(_n)
        if mode == "segments":
            gan.run_graphed = lambda name, fn: fn()
        a, b = O.synthetic_batch(1, 3, 48, seed=batch_seed)
        gan.set_input({"A": a, "B": b})
        gan.optimize_parameters()
        grads = {(n, k.replace("module.", "", 1)): p.grad.clone() for n, net in gan.networks.items() for k, p in net.named_parameters()}
        wrapped = all(isinstance(n, real_ddp) for n in gan.networks.values())
        return grads, order, (wrapped if mode == "ddp" else gan.grad_syncs is not None)

    singles = [run(False, 10 + r, "eager")[0] for r in range(world)]  # (the unsynchronised iteration is the same in both modes)
    assert any(not torch.equal(singles[0][k], singles[1][k]) for k in singles[0])  # the ranks really see different data
    out = {}
    for mode in ("eager", "segments") + (("ddp",) if recipe == "cyclegan" else ()):
        synced, order, has_sync = run(True, 10 + rank, mode)
        assert has_sync and order == (["D", "G", "mlp"] if recipe == "cut" else ["G", "D"]), (order, has_sync)
        err = 0.0
        for key, g in synced.items():
            ref = sum(s[key] for s in singles) / world
            err = max(err, float((g - ref).abs().max()) / max(1e-12, float(ref.abs().max())))
        out[mode] = err
    q.put((rank, out["eager"], out["segments"], out.get("ddp", 0.0)))
    comm.synchronize()
    dist.destroy_process_group()


def _graph_path_guarded(rank, world, port, q, recipe):
    try:
        _graph_path_worker(rank, world, port, q, recipe)
    except BaseException:  # report instead of leaving the parent waiting for its queue time-out
        import traceback
        q.put((rank, "error", traceback.format_exc()[-3000:], 0.0))
        raise


@pytest.mark.timeout(600)
@pytest.mark.parametrize("recipe", ["cyclegan", "pix2pix", "cut"])
def test_gloo_world2_graph_path_averages_gradients(recipe):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_graph_path_guarded, args=(r, world, port, q, recipe)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in range(world)), key=lambda r: r[0])
    for rank, e_eager, e_seg, e_ddp in res:
        assert e_eager != "error", e_seg
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, e_eager, e_seg, e_ddp in res:
        assert e_eager != "error", e_seg
        assert e_eager < 1e-5 and e_seg < 1e-5 and e_ddp < 1e-5, res
