"""Multi-GPU plumbing of the frame-parallel path: one process per GPU, every rank owns its own
frames (``DistributedSampler`` semantics, tasks/pmf/trainer.py:149-153), the only exchange is DDP's gradient
all-reduce over NCCL (trainer.py:38-39).  No data-path collective exists, so there is nothing to fuse."""
import os

import torch
import torch.distributed as dist


def world_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_seed(base_seed, rank):
    """Distinct synthetic frames per rank (weak scaling: fixed frames per GPU)."""
    return int(base_seed) * 1000 + int(rank)


def shard_indices(n_items, rank, world):
    """Indices of the items rank `rank` owns out of n_items, DistributedSampler-style (strided, padded by wrap-around
    so every rank gets the same count)."""
    per = (n_items + world - 1) // world
    idx = list(range(n_items)) + list(range(per * world - n_items))
    return idx[rank:per * world:world]


def wrap_ddp(model, local_rank):
    """DistributedDataParallel exactly as the reference constructs it (trainer.py:38-39)."""
    if torch.cuda.is_available() and next(model.parameters()).is_cuda:
        return torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])
    return torch.nn.parallel.DistributedDataParallel(model)


def max_over_ranks(value, device):
    """Timing rule: a multi-GPU number is the max over ranks."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
