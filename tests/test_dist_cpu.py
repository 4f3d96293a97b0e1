"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard ranges, spectral-norm replay
bookkeeping and the event-shard gather that bench.py / dist.py run over NCCL on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from v2ce_toolbox_b200 import dist as vdist
from v2ce_toolbox_b200.ldati import EVENT_DTYPE


def test_shard_ranges_cover_everything():
    for n in (1, 5, 20, 38, 563):
        for world in (1, 2, 4, 8):
            spans = [vdist.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
    assert vdist.model_calls_before(3, 'center') == 3
    assert vdist.model_calls_before(3, 'pano', tiles=2) == 6


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(rank)
        n = [1000, 37][rank]                          # ragged shards, contiguous time ranges per rank
        ev = np.zeros(n, dtype=EVENT_DTYPE)
        ev['timestamp'] = np.sort(rng.integers(rank * 10 ** 6, (rank + 1) * 10 ** 6, n))
        ev['x'] = rng.integers(0, 346, n)
        ev['y'] = rng.integers(0, 260, n)
        ev['polarity'] = rng.integers(0, 2, n)
        np.save(os.path.join(out_dir, f'shard{rank}.npy'), ev)
        t = torch.from_numpy(ev.view(np.uint8).copy())
        merged, counts = vdist.gather_event_shards(t, n)
        assert counts == [1000, 37]
        if rank == 0:
            np.save(os.path.join(out_dir, 'merged.npy'), merged.numpy().view(EVENT_DTYPE))
        else:
            assert merged is None
        # empty shard on one rank
        merged, counts = vdist.gather_event_shards(t, n if rank == 0 else 0)
        if rank == 0:
            assert merged.numel() == 1000 * 13 and counts == [1000, 0]
    finally:
        dist.destroy_process_group()


def test_gather_event_shards_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    merged = np.load(tmp_path / 'merged.npy')
    want = np.concatenate([np.load(tmp_path / f'shard{r}.npy') for r in range(world)])
    assert merged.dtype.itemsize == 13 and np.array_equal(merged, want)
    assert (np.diff(merged['timestamp']) >= 0).all()       # concatenation by rank is the time-ordered merge
