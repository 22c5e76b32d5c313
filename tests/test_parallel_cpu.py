"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: stream sharding, the filter-table broadcast and the
config-4 bus reduction.  The data path itself has no collective, so these cover everything the N>1 path adds."""
import os
import socket
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap  # noqa: E402  (spawned workers re-import this module without conftest)

_bootstrap.load_package()

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from open_headstage_b200 import parallel as P


def test_shard_range_partitions_exactly():
    for n in (1, 7, 1024, 4096, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            ranges = [P.shard_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert P.shard_range(65536, 3, 8) == (24576, 32768)   # config 3: 8192 streams per GPU
    with pytest.raises(ValueError):
        P.shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # filter table: only the source rank holds the real spectra before the broadcast
        g = torch.Generator().manual_seed(5)
        truth = torch.randn(4096, generator=g)
        table = truth.clone() if rank == 0 else torch.zeros(4096)
        P.broadcast_table(table, src=0)
        assert torch.equal(table, truth)
        # stream shards render independently: emulate "render" as a per-stream function and check the gathered result
        n_streams = 37
        lo, hi = P.shard_range(n_streams, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float32) * 2.0 + 1.0
        gathered = [torch.zeros(P.shard_range(n_streams, r, world)[1] - P.shard_range(n_streams, r, world)[0]) for r in range(world)]
        if world > 1:
            # all_gather needs equal sizes: pad to the largest shard
            m = max(t.numel() for t in gathered)
            pad = torch.zeros(m); pad[: local.numel()] = local
            bufs = [torch.zeros(m) for _ in range(world)]
            dist.all_gather(bufs, pad)
            gathered = [bufs[r][: gathered[r].numel()] for r in range(world)]
        full = torch.cat(gathered)
        assert torch.equal(full, torch.arange(n_streams, dtype=torch.float32) * 2.0 + 1.0)
        # config 4: partial stereo buses summed onto rank 0
        bus = torch.full((2, 1000), float(rank + 1))
        P.reduce_bus(bus, dst=0)
        if rank == 0:
            assert torch.equal(bus, torch.full((2, 1000), float(sum(range(1, world + 1)))))
        # partition counts travel with the table: the source derives them per set from its impulse responses,
        # every receiver marks its sets with what it received — and refuses a zero
        class FakeEngine:
            def __init__(self, parts):
                self.parts, self.marked = parts, {}

            def num_partitions(self, path, hrir_set):
                return self.parts[hrir_set][path]

            def mark_filters_external(self, hrir_set, partitions):
                self.marked[hrir_set] = partitions

        eng = FakeEngine({0: [4, 4, 3, 1], 1: [1, 1, 1, 1], 2: [47, 47, 47, 47]} if rank == 0 else {0: [1] * 4, 1: [1] * 4, 2: [1] * 4})
        counts = torch.tensor(P.set_partition_counts(eng, [0, 1, 2]) if rank == 0 else [0, 0, 0], dtype=torch.int32)
        P.broadcast_table(counts, src=0)
        P.apply_received_partition_counts(eng, [0, 1, 2], counts.tolist(), is_src=(rank == 0))
        assert eng.marked == ({} if rank == 0 else {0: 4, 1: 1, 2: 47})
        if rank != 0:
            try:
                P.apply_received_partition_counts(eng, [0], [0], is_src=False)
                raise AssertionError("a zero partition count must be refused")
            except RuntimeError:
                pass
        np.save(os.path.join(out_dir, "ok_%d.npy" % rank), np.array([1]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_broadcast_shard_reduce(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert os.path.exists(tmp_path / ("ok_%d.npy" % r))
