"""Data-parallel plumbing for the benchmark / evaluation drivers: one process per GPU, the batch sharded across ranks,
no data-path collective (images are independent).  torch.distributed is used for the barrier and the
max-over-ranks of device timings only.  Backend: nccl on GPUs, gloo in the CPU tests."""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend, device=None):
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def barrier(device=None):
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (device timings: the slowest rank defines the step time)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_range(n_items, rank, world):
    """contiguous shard [lo, hi) of n_items for this rank; sizes differ by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def whole_job_throughput(units_per_rank, world, steps, max_ms_total):
    return world * units_per_rank * steps / (max_ms_total / 1e3)


def shutdown():
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
