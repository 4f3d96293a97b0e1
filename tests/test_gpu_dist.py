"""2-GPU (NCCL) check of the sharded clip path: windows (and the width tiles of a pano window) sharded over
ranks with the spectral-norm replay, event shards gathered to rank 0 == the single-process event stream,
bit for bit.  Skips on a box with fewer than 2 GPUs (run it with ``gpurun --gpus 2``)."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import synth
from oracle.ref_harness import FakeVideoReader

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _model(seed, dev):
    from v2ce_toolbox_b200.scripts.v2ce_3d import V2ce3d
    m = V2ce3d()
    m.load_state_dict(synth.make_state_dict(seed, 'lively'))
    return m.eval().to(dev)


CASES = {
    # name: (n_frames, H, W of the source video, infer_type, model width, height, batch size)
    'center': (70, 28, 36, 'center', 36, 28, 1),           # 5 windows (last pulled back), ragged over 2 ranks
    'center_b2': (100, 28, 36, 'center', 36, 28, 2),
    'pano': (40, 24, 80, 'pano', 32, 24, 1),                # 3 width tiles per window, 3 windows
    # fewer windows than ranks: the 3 tiles of the one window are shared between the ranks (dist.pano_tile_owner)
    'pano_tiles': (17, 24, 80, 'pano', 32, 24, 1),
}


def _worker(rank, world, port, out_dir, case):
    import torch.distributed as dist
    from v2ce_toolbox_b200 import dist as vdist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        n_frames, H, W, infer_type, width, height, bs = CASES[case]
        frames = synth.make_video(n_frames, H, W, seed=5)
        ev, n = vdist.stream_clip_sharded(_model(31, dev), FakeVideoReader(frames), n_frames, world, rank, seq_len=16,
                                          batch_size=bs, infer_type=infer_type, width=width, height=height, fps=30, seed=9,
                                          device=dev)
        if rank == 0:
            np.save(os.path.join(out_dir, f'{case}.npy'), ev)
            assert len(ev) == n
        else:
            assert ev is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('case', list(CASES))
def test_sharded_clip_equals_single_process(case, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    from v2ce_toolbox_b200 import v2ce as drv
    n_frames, H, W, infer_type, width, height, bs = CASES[case]
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), case), nprocs=2, join=True)
    got = np.load(tmp_path / f'{case}.npy')
    frames = synth.make_video(n_frames, H, W, seed=5)
    res = drv.stream_clip(_model(31, 'cuda:0'), vidcap=FakeVideoReader(frames), infer_type=infer_type, seq_len=16,
                          width=width, height=height, batch_size=bs, fps=30, seed=9, write_event_frames=False)
    want = res.event_stream
    assert got.dtype.itemsize == 13 and len(got) == len(want) and len(want) > 0
    for f in ('timestamp', 'x', 'y', 'polarity'):
        assert np.array_equal(got[f], want[f]), (case, f)


def _preview_worker(rank, world, port, out_dir, case):
    import torch.distributed as dist
    from v2ce_toolbox_b200 import dist as vdist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        n_frames, H, W, infer_type, width, height, bs = CASES[case]
        frames = synth.make_video(n_frames, H, W, seed=5)
        preview = {}
        ev, n = vdist.stream_clip_sharded(_model(31, dev), FakeVideoReader(frames), n_frames, world, rank, seq_len=16,
                                          batch_size=bs, infer_type=infer_type, width=width, height=height, fps=30, seed=9,
                                          device=dev, preview=preview, ceil=10, upper_bound_percentile=98)
        if rank == 0:
            np.save(os.path.join(out_dir, f'{case}_frames.npy'), preview['frames'])
            np.save(os.path.join(out_dir, f'{case}_ub.npy'), np.float64(preview['upper_bound']))
            np.save(os.path.join(out_dir, f'{case}_events.npy'), ev)
        else:
            assert 'frames' not in preview
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('case,world', [('center', 1), ('center', 2), ('pano', 2), ('pano_tiles', 2)])
def test_sharded_preview_equals_single_process(case, world, tmp_path):
    """the event-frame preview of a sharded clip -- per-rank sums gathered to rank 0, ONE clip-global
    percentile -- is the single-process preview, byte for byte (world 1 runs the same code on one GPU)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    import torch.multiprocessing as mp
    from v2ce_toolbox_b200 import v2ce as drv
    n_frames, H, W, infer_type, width, height, bs = CASES[case]
    mp.spawn(_preview_worker, args=(world, _free_port(), str(tmp_path), case), nprocs=world, join=True)
    frames = synth.make_video(n_frames, H, W, seed=5)
    res = drv.stream_clip(_model(31, 'cuda:0'), vidcap=FakeVideoReader(frames), infer_type=infer_type, seq_len=16,
                          width=width, height=height, batch_size=bs, fps=30, seed=9, write_event_frames=True,
                          ceil=10, upper_bound_percentile=98)
    assert float(np.load(tmp_path / f'{case}_ub.npy')) == res.ef_upper_bound
    assert np.array_equal(np.load(tmp_path / f'{case}_frames.npy'), res.ef_frames)
    assert np.array_equal(np.load(tmp_path / f'{case}_events.npy').view(np.uint8), res.event_stream.view(np.uint8))


def _window_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from v2ce_toolbox_b200 import dist as vdist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    vdist.init_process_group('nccl', device=dev, rank=rank, world_size=world)
    try:
        win = vdist.PeerWindow(4 << 20, buffers=2, device=dev)
        for step, n in enumerate([1000 + 37 * rank, 0 if rank == 0 else 5, 50000 - 9 * rank, 0]):
            g = torch.Generator().manual_seed(100 * step + rank)
            shard = torch.randint(0, 256, (n * 13 + 64,), dtype=torch.uint8, generator=g).to(dev)
            via_nccl, counts = vdist.gather_event_shards(shard, n)
            via_win, counts2 = vdist.gather_event_shards(shard, n, window=win, slot=step & 1)
            assert counts == counts2
            if rank == 0:
                assert via_win.numel() == sum(counts) * 13
                assert torch.equal(via_win, via_nccl), step
                np.save(os.path.join(out_dir, f'win{step}.npy'), via_win.cpu().numpy())
            else:
                assert via_win is None
        win.close()
    finally:
        vdist.close_host_group()
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [1, 2])
def test_peer_window_gather_equals_nccl_gather(world, tmp_path):
    """dist.PeerWindow (CUDA IPC window on rank 0, one copy-engine push per rank) merges ragged shards -- including
    empty ones -- exactly like the NCCL point-to-point gather; two alternating buffers."""
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    import torch.multiprocessing as mp
    mp.spawn(_window_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for step, n in enumerate([1000, 0, 50000, 0]):
        want = []
        for rank in range(world):
            k = [1000 + 37 * rank, 0 if rank == 0 else 5, 50000 - 9 * rank, 0][step]
            g = torch.Generator().manual_seed(100 * step + rank)
            want.append(torch.randint(0, 256, (k * 13 + 64,), dtype=torch.uint8, generator=g)[:k * 13].numpy())
        assert np.array_equal(np.load(tmp_path / f'win{step}.npy'), np.concatenate(want))
