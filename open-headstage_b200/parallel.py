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


def set_partition_counts(engine, hrir_sets) -> list[int]:
    """Partitions in use per HRIR set: the longest of its four paths (what the render kernel walks)."""
    return [max(engine.num_partitions(p, s) for p in range(4)) for s in hrir_sets]


def broadcast_filters(engine, src: int = 0, hrir_sets=None, group=None) -> None:
    """Rank `src` must already have called set_ir for `hrir_sets` (default: every set of the engine); the other ranks
    receive the whole device-resident spectra table and, per set, the partition count rank `src` derived from its
    impulse responses."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    sets = list(range(engine.n_hrir_sets)) if hrir_sets is None else [int(s) for s in hrir_sets]
    dev = torch.device("cuda", engine.device)
    if rank == src:
        engine.commit_filters()
        engine.sync()
        counts = set_partition_counts(engine, sets)
    else:
        counts = [0] * len(sets)
    ptr, nbytes = engine.filter_table()
    with torch.cuda.device(dev):
        table = torch.as_tensor(DeviceMemory(ptr, nbytes), device=dev)
        if table.data_ptr() != ptr:
            raise RuntimeError("torch copied the engine's filter table instead of viewing it (device mismatch?)")
        parts = torch.tensor(counts, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        broadcast_table(table, src, group)
        broadcast_table(parts, src, group)
        torch.cuda.synchronize(dev)
    apply_received_partition_counts(engine, sets, parts.tolist(), rank == src)


def apply_received_partition_counts(engine, hrir_sets, counts, is_src: bool) -> None:
    """After a table broadcast: every receiving rank marks the sets as externally filled with the source's counts."""
    for s, c in zip(hrir_sets, counts):
        if c < 1:
            raise RuntimeError("broadcast_filters: received partition count %d for HRIR set %d" % (c, s))
        if not is_src:
            engine.mark_filters_external(s, int(c))


def reduce_bus(bus: torch.Tensor, dst: int = 0, group=None) -> torch.Tensor:
    """Sum the per-rank partial stereo buses [2, n_frames] onto rank `dst` (config 4)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(bus, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return bus


def render_object_mix(pkg, sources, hrirs, block: int, fs: float, eq_preset=None, gain: float = 1.0, device: int = 0,
                      src: int = 0, group=None):
    """BASELINE config 4 on this rank's share: `sources` [n_src, n_frames] mono signals (n_src even), `hrirs`
    [n_src, 2, taps] (left-ear, right-ear impulse responses of each source's direction).  Two sources ride in one stereo
    stream of the engine (source 2i on the left input with paths LSL/LSR, source 2i+1 on the right input with RSL/RSR);
    the rendered streams are summed into this rank's stereo bus on the GPU, the buses are summed onto rank `src` with
    NCCL, and the EQ and the output gain are applied ONCE to the reduced bus there (the definition of SURVEY.md §8e —
    the reference has no multi-source mode).  Returns the [2, n_frames] bus (torch, CUDA) on `src`, None elsewhere."""
    import numpy as np

    n_src, n_frames = sources.shape
    assert n_src % 2 == 0 and n_frames % block == 0 and n_frames % 4 == 0
    taps = hrirs.shape[2]
    n_streams = n_src // 2
    eng = pkg.Engine(n_streams, block, taps, n_bands=0, n_hrir_sets=n_streams, device=device, sample_rate=fs)
    for i in range(n_streams):
        eng.set_ir(0, hrirs[2 * i, 0], hrir_set=i)      # LSL: source 2i   -> left ear
        eng.set_ir(1, hrirs[2 * i, 1], hrir_set=i)      # LSR: source 2i   -> right ear
        eng.set_ir(2, hrirs[2 * i + 1, 0], hrir_set=i)  # RSL: source 2i+1 -> left ear
        eng.set_ir(3, hrirs[2 * i + 1, 1], hrir_set=i)  # RSR: source 2i+1 -> right ear
        eng.bind_stream_hrir(i, i)
    x = torch.from_numpy(np.ascontiguousarray(sources.reshape(n_streams, 2, n_frames), dtype=np.float32)).to("cuda:%d" % device)
    y = torch.empty_like(x)
    bus = torch.zeros((2, n_frames), dtype=torch.float32, device=x.device)
    torch.cuda.synchronize(x.device)
    eng.process_device(x.data_ptr(), y.data_ptr(), n_frames)
    eng.mix_device(y.data_ptr(), bus.data_ptr(), n_frames)
    eng.sync()
    reduce_bus(bus, dst=src, group=group)
    rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0
    if rank != src:
        return None
    if eq_preset is not None or gain != 1.0:
        post = pkg.Engine(1, block, 1, n_bands=10, device=device, sample_rate=fs)
        post.set_conv_enable(False)
        if eq_preset is not None:
            post.eq_set_preset(eq_preset)
            post.set_eq_enable(True)
        post.set_gain(gain)
        torch.cuda.synchronize(x.device)
        post.process_device(bus.data_ptr(), bus.data_ptr(), n_frames)
        post.sync()
    return bus
