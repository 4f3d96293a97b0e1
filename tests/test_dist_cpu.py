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


def test_shard_schedules_tile_the_clip():
    """Every frame pair of the clip is emitted by exactly one rank, in order, including the pulled-back last
    window when it is the only window of the last rank (F8: 40 frames -> starts [0,16,23], mode 7)."""
    from v2ce_toolbox_b200 import v2ce as drv
    for n_frames in (17, 18, 40, 70, 100, 321, 600):
        for bs in (1, 2, 4):
            starts, mode = drv.window_schedule(n_frames, 16)
            n_batches = -(-len(starts) // bs)
            for world in (1, 2, 3, 8):
                nxt = 0
                for r in range(world):
                    b0, b1 = vdist.shard_range(n_batches, world, r)
                    first, count, tail, pair_base, (rel, m) = vdist.shard_schedule(starts, mode, b0, b1, bs, 16)
                    if count == 0:
                        continue
                    assert pair_base == nxt and first + rel[0] == starts[b0 * bs]
                    assert count == first + rel[-1] + 17 - first and (m == mode if tail else m == 0)
                    emitted = 16 * len(rel) - ((16 - m) if (tail and m) else 0)
                    nxt += emitted
                assert nxt == n_frames - 1, (n_frames, bs, world)


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
        # event-frame sums of a sharded clip: ragged leading dimension, one rank possibly without any window
        for rows_per_rank in ([5, 3], [4, 0], [0, 2]):
            k = rows_per_rank[rank]
            base = sum(rows_per_rank[:rank])
            rows = (torch.arange(k * 2 * 3 * 4, dtype=torch.float32).reshape(k, 2, 3, 4) + 1000 * base)
            got = vdist.gather_row_shards(rows)
            if rank == 0:
                want = torch.cat([torch.arange(r * 24, dtype=torch.float32).reshape(r, 2, 3, 4) + 1000 * sum(rows_per_rank[:i])
                                  for i, r in enumerate(rows_per_rank)], dim=0)
                assert got.shape == want.shape and torch.equal(got, want)
            else:
                assert got is None
        # the shared host ring of the multi-GPU e2e leg (page-locking skipped on CPU): every rank's bytes land in its
        # slice of the rank-ordered merged array, ragged and empty shards included
        ring = vdist.SharedHostRing(2, 4096, register=False)
        for slot, sizes in ((0, [100, 37]), (1, [0, 500]), (0, [13, 0])):
            view = ring.place(slot, sizes[rank])
            view[:sizes[rank]] = rank + 1
            dist.barrier()
            assert ring.layout(slot) == [(0, sizes[0]), (sizes[0], sizes[1])]
            merged = ring.merged(slot)
            assert merged.numel() == sum(sizes)
            assert bool((merged[:sizes[0]] == 1).all()) and bool((merged[sizes[0]:] == 2).all())
            dist.barrier()
        with pytest.raises(ValueError):
            ring.place(0, 5000)
        ring.close()
    finally:
        dist.destroy_process_group()


def test_gather_event_shards_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    merged = np.load(tmp_path / 'merged.npy')
    want = np.concatenate([np.load(tmp_path / f'shard{r}.npy') for r in range(world)])
    assert merged.dtype.itemsize == 13 and np.array_equal(merged, want)
    assert (np.diff(merged['timestamp']) >= 0).all()       # concatenation by rank is the time-ordered merge


def test_bind_to_gpu_numa_is_a_no_op_without_nvml_device():
    """bench.py calls it on every rank at N > 1; on a host without a usable NVML device it must change nothing."""
    import os
    from v2ce_toolbox_b200 import dist as vdist
    before = os.sched_getaffinity(0)
    assert vdist.bind_to_gpu_numa(0) is None or isinstance(vdist.bind_to_gpu_numa(0), list)
    import torch
    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before


def test_pano_tile_groups():
    """dist.pano_tile_owner: fewer batches than ranks -> groups of world // n_batches ranks share a window's tiles; the
    first member of a group reports the batch; ranks beyond the last group idle; otherwise the window split applies."""
    from v2ce_toolbox_b200.dist import pano_tile_owner
    assert pano_tile_owner(1, 2, 0) == (0, [0, 1], True)
    assert pano_tile_owner(1, 2, 1) == (0, [0, 1], False)
    assert pano_tile_owner(2, 8, 5) == (1, [4, 5, 6, 7], False)
    assert pano_tile_owner(3, 8, 4) == (2, [4, 5], True)
    assert pano_tile_owner(3, 8, 7) == (None, [], False)
    assert pano_tile_owner(5, 8, 0) is None and pano_tile_owner(8, 8, 3) is None and pano_tile_owner(0, 4, 0) is None
    for n_batches, world in [(1, 2), (1, 8), (2, 8), (3, 8), (2, 4)]:
        leaders = [r for r in range(world) if (pano_tile_owner(n_batches, world, r) or (None,))[0] is not None
                   and pano_tile_owner(n_batches, world, r)[2]]
        assert [pano_tile_owner(n_batches, world, r)[0] for r in leaders] == list(range(n_batches))
