"""Multi-GPU plumbing.  Streams are independent, so the data path shards with NO collective: rank r of
W renders the contiguous stream range shard_range(n, r, W).  torch.distributed (NCCL over NVLink on GPUs, gloo in the
CPU tests) is used for exactly two things:

  * broadcast_filters  — one HRIR spectra table for the whole job: rank `src` transforms the impulse responses once,
    everyone else receives the device-resident table bit for bit (instead of each rank recomputing it);
  * reduce_bus         — BASELINE config 4 (object mixdown): every rank mixes its sources into a partial stereo bus,
    the buses are summed onto `dst`, once per render chunk (a 10 s chunk is 3.84 MB; per-block reduces would be
    latency-bound 2 KB messages).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) of `n_items` for `rank` (the first n_items % world ranks get one extra)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class DeviceMemory:
    """Wraps a raw device pointer (e.g. ohs_filter_table) so torch can view it without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def broadcast_table(table: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    """In-place broadcast of a filter-spectra table (any device; NCCL for CUDA tensors, gloo for CPU tensors)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(table, src=src, group=group)
    return table


def broadcast_filters(engine, src: int = 0, partitions: int | None = None, hrir_sets=(0,), group=None) -> None:
    """Rank `src` must already have called set_ir for `hrir_sets`; the other ranks receive the spectra."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if rank == src:
        engine.commit_filters()
        engine.sync()
    ptr, nbytes = engine.filter_table()
    table = torch.as_tensor(DeviceMemory(ptr, nbytes), device="cuda")
    torch.cuda.synchronize()
    parts = torch.tensor([partitions or 0], dtype=torch.int32, device="cuda")
    broadcast_table(table, src, group)
    broadcast_table(parts, src, group)
    torch.cuda.synchronize()
    if rank != src:
        for s in hrir_sets:
            engine.mark_filters_external(s, int(parts.item()) or 1)


def reduce_bus(bus: torch.Tensor, dst: int = 0, group=None) -> torch.Tensor:
    """Sum the per-rank partial stereo buses [2, n_frames] onto rank `dst` (config 4)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(bus, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return bus
