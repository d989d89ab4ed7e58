"""Multi-GPU sharding of a simulation batch: one process per GPU, datasets partitioned into
contiguous ranges, no data-path collective.  The Philox counter carries the GLOBAL dataset
index, so the assembled batch is bit-identical for any world size.
The only collective is the optional all-gather that assembles a training batch.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous range [lo, hi) of rank's share of n items (sizes differ by at most 1)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank outside [0, world_size)")
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def world():
    """(rank, world_size) from torch.distributed if initialised, else env, else (0, 1)."""
    import os

    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


_batch_base = 0  # global dataset index of the next sharded batch (advances by B on every rank alike)


def simulate_sharded(simulate_fn, params, n_trials, *, rank=None, world_size=None, dataset_base=None, gather=False,
                     **kw):
    """Simulate this rank's slice of a (B, P) parameter batch.

    ``dataset_base`` is the global index of the batch's first dataset.  ``None`` (default) takes it from a
    process-wide counter that advances by B per call -- every rank calls with the same B, so the ranks stay
    in step and successive batches use fresh Philox counters; pass an explicit base to regenerate a batch.

    ``simulate_fn(params_local, n_trials, dataset_offset=..., **kw)`` is a model module's
    ``batch_simulate_trials`` (numpy out) or ``batch_simulate_trials_device`` (DLPack out).
    Returns (local_batch, (lo, hi)); with ``gather=True`` the first element is the full
    (B, n_trials, 2) batch on every rank (torch.distributed all-gather: NCCL for device
    tensors, gloo for host arrays)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    params = np.asarray(params, dtype=np.float64)
    lo, hi = shard_range(params.shape[0], rank, world_size)
    if dataset_base is None:
        global _batch_base
        B = int(params.shape[0])
        if (_batch_base & 0xFFFFFFFF) + B > 1 << 32:   # a launch may not straddle a multiple of 2^32
            _batch_base = ((_batch_base >> 32) + 1) << 32
        dataset_base = _batch_base
        _batch_base += B
    local = simulate_fn(params[lo:hi], n_trials, dataset_offset=int(dataset_base) + lo, **kw)
    if not gather or world_size == 1:
        return local, (lo, hi)
    return all_gather_batch(local, params.shape[0], world_size), (lo, hi)


def all_gather_batch(local, n_total: int, world_size: int):
    """Assemble per-rank (B_r, N, D) shards into (n_total, N, D) on every rank."""
    import torch
    import torch.distributed as dist

    was_numpy = isinstance(local, np.ndarray)
    t = torch.from_numpy(local) if was_numpy else (local if isinstance(local, torch.Tensor) else torch.from_dlpack(local))
    sizes = [shard_range(n_total, r, world_size) for r in range(world_size)]
    max_rows = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == max_rows for lo, hi in sizes):
        full = torch.empty((n_total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, t.contiguous())
    else:
        pad = torch.zeros((max_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world_size)]
        dist.all_gather(parts, pad)
        full = torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
    return full.numpy() if was_numpy else full
