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


def create_comm(pkg, device: int, group=None):
    """An ohs_comm (NCCL communicator owned by the C ABI library) over the ranks of the torch.distributed job: rank 0
    asks NCCL for the unique id through the C ABI, torch.distributed only ships the 128 bytes."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ids = [pkg.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(ids, src=0, group=group)
    return pkg.Comm(world, rank, device, ids[0])


def broadcast_filters(engine, src: int = 0, hrir_sets=None, group=None, comm=None) -> None:
    """Rank `src` must already have called set_ir for `hrir_sets` (default: every set of the engine); the other ranks
    receive the whole device-resident spectra table and, per set, the partition count rank `src` derived from its
    impulse responses.  With `comm` (create_comm) the whole exchange is the C ABI's ohs_broadcast_hrir; without it the
    table is viewed as a torch tensor and torch.distributed broadcasts it."""
    if comm is not None:
        engine.broadcast_hrir(comm, src)
        return
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


class ObjectMixer:
    """BASELINE config 4 on this rank's share.  `hrirs` [n_src, 2, taps] (left-ear, right-ear impulse responses of each
    source's direction, n_src even).  Two sources ride in one stereo stream of the engine (source 2i on the left input
    with paths LSL/LSR, source 2i+1 on the right input with RSL/RSR); render() sums the rendered streams into this
    rank's stereo bus on the GPU, the buses are summed onto rank `dst` with NCCL, and the EQ and the output gain are
    applied ONCE to the reduced bus there (the definition of SURVEY.md §8e — the reference has no multi-source mode).

    All set-up (engine, one HRIR set per stream, filter transforms, scratch) happens in the constructor; render() holds
    only the data path, pipelined over time chunks: process(c) -> mix(c) -> reduce(c) on the engine's stream while the
    bus EQ of chunk c-1 — one strictly sequential biquad chain, the longest single dependency of this config — runs on
    its own stream."""

    def __init__(self, pkg, hrirs, block: int, fs: float, n_frames: int, eq_preset=None, gain: float = 1.0, device: int = 0,
                 dst: int = 0, group=None, comm=None, chunk_blocks: int = 96):
        import numpy as np

        self.comm = comm   # create_comm(): the bus reduce is then the C ABI's ohs_reduce_bus on the engine's stream

        n_src, two, taps = hrirs.shape
        assert two == 2 and n_src % 2 == 0 and n_frames % block == 0 and n_frames % 4 == 0
        self.n_src, self.n_streams, self.n_frames, self.block, self.dst, self.group = n_src, n_src // 2, n_frames, block, dst, group
        self.chunk = min(n_frames, chunk_blocks * block)
        self.device = torch.device("cuda", device)
        hrirs = np.ascontiguousarray(hrirs, dtype=np.float32)
        self.engine = eng = pkg.Engine(self.n_streams, block, taps, n_bands=0, n_hrir_sets=self.n_streams, device=device, sample_rate=fs)
        for i in range(self.n_streams):
            eng.set_ir(0, hrirs[2 * i, 0], hrir_set=i)      # LSL: source 2i   -> left ear
            eng.set_ir(1, hrirs[2 * i, 1], hrir_set=i)      # LSR: source 2i   -> right ear
            eng.set_ir(2, hrirs[2 * i + 1, 0], hrir_set=i)  # RSL: source 2i+1 -> left ear
            eng.set_ir(3, hrirs[2 * i + 1, 1], hrir_set=i)  # RSR: source 2i+1 -> right ear
            eng.bind_stream_hrir(i, i)
        eng.prepare(self.chunk)                              # one upload + one transform launch for all the sets
        self.rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0
        self.post = None
        if self.rank == dst and (eq_preset is not None or gain != 1.0):
            # EQ-only engine for the bus; block 1024 selects the one-band-per-lane EQ warps (the shorter step of the two)
            self.post = post = pkg.Engine(1, 1024, 1, n_bands=10, device=device, sample_rate=fs)
            post.set_conv_enable(False)
            if eq_preset is not None:
                post.eq_set_preset(eq_preset)
                post.set_eq_enable(True)
            post.set_gain(gain)
            post.prepare(self.chunk)
        with torch.cuda.device(self.device):
            self.rendered = torch.empty((self.n_streams, 2, n_frames), dtype=torch.float32, device=self.device)
            self.bus = torch.zeros((2, n_frames), dtype=torch.float32, device=self.device)
            self._eng_stream = torch.cuda.ExternalStream(eng.cuda_stream(), device=self.device)
            self._post_stream = torch.cuda.ExternalStream(self.post.cuda_stream(), device=self.device) if self.post else None
            self._red_stream = torch.cuda.Stream(self.device)   # torch.distributed's reduce runs here when no ohs_comm is given
        eng.sync()

    def reset(self):
        self.engine.conv_reset()
        if self.post:
            self.post.eq_reset()

    def render(self, sources: torch.Tensor, apply_post: bool = True):
        """sources [n_src, n_frames] f32 on this rank's GPU (viewed as [n_streams, 2, n_frames]).  Returns the reduced,
        equalised [2, n_frames] bus on rank `dst` (valid after torch.cuda.synchronize()), None elsewhere.
        apply_post=False leaves the reduced bus raw (no EQ, no gain)."""
        assert sources.is_cuda and sources.dtype == torch.float32 and sources.is_contiguous() and tuple(sources.shape) == (self.n_src, self.n_frames)
        eng = self.engine
        cur = torch.cuda.current_stream(self.device)
        n, fsz = self.n_frames, 4
        self._eng_stream.wait_stream(cur)
        red = self._red_stream
        post_on = self.post is not None and apply_post and self.rank == self.dst
        for c0 in range(0, n, self.chunk):
            m = min(self.chunk, n - c0)
            eng.process_device(sources.data_ptr() + c0 * fsz, self.rendered.data_ptr() + c0 * fsz, m, row_stride=n)
            eng.mix_device(self.rendered.data_ptr() + c0 * fsz, self.bus.data_ptr() + c0 * fsz, m, row_stride=n, bus_stride=n)
            if self.comm is not None:
                for ch in range(2):   # NCCL, on the engine's stream
                    eng.reduce_bus(self.comm, self.bus.data_ptr() + (ch * n + c0) * fsz, m, self.dst)
                done = self._eng_stream.record_event()
            else:
                red.wait_stream(self._eng_stream)
                with torch.cuda.stream(red):
                    for ch in range(2):
                        reduce_bus(self.bus[ch, c0:c0 + m], dst=self.dst, group=self.group)
                done = red.record_event()
            if post_on:
                self._post_stream.wait_event(done)
                self.post.process_device(self.bus.data_ptr() + c0 * fsz, self.bus.data_ptr() + c0 * fsz, m, row_stride=n)
            else:
                cur.wait_event(done)
        if post_on:
            cur.wait_stream(self._post_stream)
        cur.wait_stream(self._eng_stream)
        return self.bus if self.rank == self.dst else None


def render_object_mix(pkg, sources, hrirs, block: int, fs: float, eq_preset=None, gain: float = 1.0, device: int = 0,
                      src: int = 0, group=None):
    """One-shot form of ObjectMixer: `sources` [n_src, n_frames] mono signals on the host.  Returns the [2, n_frames] bus
    (torch, CUDA) on rank `src`, None elsewhere."""
    import numpy as np

    n_src, n_frames = sources.shape
    mixer = ObjectMixer(pkg, hrirs, block, fs, n_frames, eq_preset=eq_preset, gain=gain, device=device, dst=src, group=group)
    x = torch.from_numpy(np.ascontiguousarray(sources, dtype=np.float32)).to(mixer.device)
    torch.cuda.synchronize(mixer.device)
    bus = mixer.render(x)
    torch.cuda.synchronize(mixer.device)
    return bus.clone() if bus is not None else None
