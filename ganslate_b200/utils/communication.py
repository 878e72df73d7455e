"""Distributed helpers with the semantics of ganslate/utils/communication.py:17-284 (NCCL, env://, one process
per GPU) -- the parts the training hot path uses."""
import os

import numpy as np
import torch
import torch.distributed as dist


def init_distributed():
    """communication.py:17-27: WORLD_SIZE>1 => NCCL process group from the environment + barrier."""
    num_gpu = int(os.environ.get("WORLD_SIZE", 1))
    if num_gpu > 1 and not dist.is_initialized():
        backend = "nccl" if torch.cuda.is_available() else "gloo"  # gloo only for the CPU host-logic tests
        if backend == "nccl":
            torch.cuda.set_device(get_local_rank())
            dist.init_process_group(backend=backend, init_method="env://",
                                    device_id=torch.device("cuda", get_local_rank()))
        else:
            dist.init_process_group(backend=backend, init_method="env://")
        synchronize()


def synchronize():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def get_rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_local_rank() -> int:
    if not (dist.is_available() and dist.is_initialized()):
        return int(os.environ.get("LOCAL_RANK", 0)) if "LOCAL_RANK" in os.environ else 0
    return int(os.environ["LOCAL_RANK"])


def get_world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def is_main_process() -> bool:
    return get_rank() == 0


def shared_random_seed() -> int:
    """communication.py:101-116: a seed every rank agrees on (broadcast from rank 0)."""
    seed = np.random.randint(2**31)
    if get_world_size() == 1:
        return seed
    dev = torch.device("cuda") if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([seed], dtype=torch.int64, device=dev)
    dist.broadcast(t, src=0)
    return int(t.item())


def reduce(input_data, average=False, all_reduce=False):
    """communication.py:153-284: reduce a tensor / dict of scalars to rank 0 (or all ranks)."""
    if get_world_size() < 2:
        return input_data
    if isinstance(input_data, dict):
        keys = sorted(input_data)
        vals = torch.stack([torch.as_tensor(input_data[k], dtype=torch.float32).detach().reshape(()) for k in keys])
        vals = reduce(vals.to(_comm_device()), average, all_reduce)
        return {k: v for k, v in zip(keys, vals)}
    t = input_data.detach().clone().to(_comm_device())
    if all_reduce:
        dist.all_reduce(t)
    else:
        dist.reduce(t, dst=0)
    if average and (all_reduce or get_rank() == 0):
        t = t / get_world_size()
    return t


def _comm_device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
