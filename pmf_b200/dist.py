"""Multi-GPU plumbing of the frame-parallel path: one process per GPU, every rank owns its own
frames (``DistributedSampler`` semantics, tasks/pmf/trainer.py:149-153), the only exchange is the gradient all-reduce
(trainer.py:38-39) — 36.4 M fp32 = 146 MB per step over NCCL / NVLink.  No data-path collective exists.

Two ways to run it:

* stock ``torch.nn.parallel.DistributedDataParallel`` around ``pmf_b200.PMFNet`` — what the unchanged trainer does.  It
  works (tests/test_dropin_gpu.py) but pays DDP's per-parameter bookkeeping every step: 372 gradient copies into the
  buckets, bucket views, the reducer's finalisation — ~3.5 ms per step measured on 2 x B200, of which the NCCL transfer
  itself is ~0.4 ms.
* ``FrameParallel`` (below): the module's backward already writes every parameter gradient into ONE flat buffer, segment
  by segment (pmf_b200.modules._GraphedPMF); each segment's slice is all-reduced (ReduceOp.AVG) asynchronously as soon as
  its CUDA graph has been replayed, overlapping the next segments, and autograd receives views of the reduced slice.
  No per-parameter copies, no buckets.  Same semantics as DDP: parameters broadcast from rank 0 at construction, BN
  running-statistic buffers broadcast from rank 0 before every training forward (DDP's ``broadcast_buffers=True`` default,
  which the reference does not override), averaged gradients, ``no_sync()`` for gradient accumulation.
"""
import contextlib
import os

import torch
import torch.distributed as dist
import torch.nn as nn


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


def _broadcast_flat(tensors, group):
    """Broadcast a list of same-dtype tensors from rank 0 with ONE collective (flatten -> broadcast -> multi-tensor copy)."""
    if not tensors:
        return
    flat = torch.cat([t.detach().reshape(-1) for t in tensors])
    dist.broadcast(flat, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    outs = torch.split(flat, [t.numel() for t in tensors])
    torch._foreach_copy_([t.detach() for t in tensors], [o.view_as(t) for o, t in zip(outs, tensors)])


class GradSync:
    """Asynchronous averaging of the flat gradient slices a PMFNet backward produces (see FrameParallel)."""

    def __init__(self, group=None):
        self.group = group
        self.enabled = True
        self.pending = []
        self._cb_queued = False

    def world(self):
        return dist.get_world_size(self.group)

    def _all_reduce_avg(self, t, async_op=False):
        """Average over the ranks.  NCCL averages inside the collective; other backends (gloo, CPU tests) sum and the
        division follows (after the wait, for asynchronous calls)."""
        if dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op), None
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if not async_op:
            t.div_(self.world())
            return w, None
        return w, t

    def reduce(self, flat_slice, params):
        """All-reduce (AVG) ``flat_slice`` in place; returns immediately.  The compute stream is made to wait for every
        pending reduction at the end of the backward pass (autograd callback), before any optimiser can read .grad."""
        if not self.enabled or self.world() == 1:
            return
        if any(p.grad is not None for p in params):
            # gradient accumulation into an existing .grad reads the slice on the compute stream right away
            self._all_reduce_avg(flat_slice)
            return
        self.pending.append(self._all_reduce_avg(flat_slice, async_op=True))
        if not self._cb_queued:
            self._cb_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self.finish)

    def reduce_list(self, tensors):
        """Synchronous average of separate gradient tensors through one flat buffer (the eager, un-captured pass)."""
        if not tensors or not self.enabled or self.world() == 1:
            return
        flat = torch.cat([t.reshape(-1) for t in tensors])
        self._all_reduce_avg(flat)
        torch._foreach_copy_(list(tensors), [o.view_as(t) for o, t in zip(torch.split(flat, [t.numel() for t in tensors]), tensors)])

    def finish(self):
        for w, t in self.pending:
            w.wait()
            if t is not None:
                t.div_(self.world())
        self.pending = []
        self._cb_queued = False


class FrameParallel(nn.Module):
    """DistributedDataParallel semantics for pmf_b200.PMFNet with the gradient all-reduce issued per backward segment on
    the module's own flat gradient buffer (no buckets, no per-parameter copies)."""

    def __init__(self, module, broadcast_buffers=True, process_group=None):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("FrameParallel needs an initialised torch.distributed process group")
        self.module = module
        self.group = process_group
        self.broadcast_buffers = broadcast_buffers
        self.sync = GradSync(process_group)
        module._grad_sync = self.sync
        with torch.no_grad():
            by_dtype = {}
            for t in list(module.parameters()) + list(module.buffers()):
                by_dtype.setdefault(t.dtype, []).append(t)
            for ts in by_dtype.values():
                _broadcast_flat(ts, self.group)

    def forward(self, *args, **kwargs):
        if self.broadcast_buffers and self.module.training and torch.is_grad_enabled() and self.sync.enabled:
            with torch.no_grad():
                by_dtype = {}
                for t in self.module.buffers():
                    by_dtype.setdefault(t.dtype, []).append(t)
                for ts in by_dtype.values():
                    _broadcast_flat(ts, self.group)
        return self.module(*args, **kwargs)

    @contextlib.contextmanager
    def no_sync(self):
        prev, self.sync.enabled = self.sync.enabled, False
        try:
            yield
        finally:
            self.sync.enabled = prev


def wrap_ddp(model, local_rank, flat=None):
    """Frame-parallel wrapper.  ``flat=None``: FrameParallel for a CUDA pmf_b200.PMFNet, else stock
    DistributedDataParallel exactly as the reference constructs it (trainer.py:38-39)."""
    is_cuda = torch.cuda.is_available() and next(model.parameters()).is_cuda
    if flat is None:
        flat = is_cuda and hasattr(model, "_graphs")
    if flat:
        return FrameParallel(model)
    if is_cuda:
        return torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])
    return torch.nn.parallel.DistributedDataParallel(model)


def max_over_ranks(value, device):
    """Timing rule: a multi-GPU number is the max over ranks."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
