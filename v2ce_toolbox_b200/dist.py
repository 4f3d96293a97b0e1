"""Multi-GPU plumbing: one process per GPU (torchrun), windows sharded by contiguous ranges, NCCL only
for the final gather of the per-rank event shards.  No data-path collective: windows,
frame pairs and pano tiles are independent; the spectral-norm schedule is input-independent, so a rank
replays (``V2ce3d.sn_advance``) the power-iteration steps of the model calls it does not own.

The same functions run on CPU tensors over the ``gloo`` backend (tests/test_dist_cpu.py).
"""
import numpy as np
import torch
import torch.distributed as dist

EVENT_BYTES = 13


def bind_to_gpu_numa(device_index):
    """Restrict this process to the CPUs NVML reports as local to CUDA device `device_index` (one process per GPU:
    its pinned staging buffers are then first-touched on the GPU's own NUMA node and the D2H / H2D copies do not cross
    the socket interconnect).  Returns the sorted CPU list, or None when NVML, the device or a usable mask is not
    available (nothing is changed then).  V2CE_NO_AFFINITY=1 disables it."""
    import os
    if os.environ.get('V2CE_NO_AFFINITY', '0') not in ('', '0') or not hasattr(os, 'sched_setaffinity'):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid).encode())
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, (max(os.cpu_count() or 1, 1) + 63) // 64)
        finally:
            pynvml.nvmlShutdown()
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if len(cpus) < 8:                     # a mask this small would crowd the rank's copy and decode threads
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def shard_range(n_items, world, rank):
    """Contiguous [start, end) of `n_items` owned by `rank`: rank r gets items [r*ceil(n/R), ...)."""
    per = -(-n_items // world)
    start = min(rank * per, n_items)
    return start, min(start + per, n_items)


def model_calls_before(first_batch, infer_type='center', tiles=1):
    """Number of V2ce3d forwards the single-process schedule has executed before batch `first_batch`
    (v2ce.py:181-198: one call per batch in center mode, one per tile per batch in pano mode)."""
    return first_batch * (1 if infer_type == 'center' else tiles)


def gather_event_shards(events_u8, n_events, group=None, dst=0):
    """events_u8: 1-D uint8 tensor holding n_events*13 bytes (device for NCCL, CPU for gloo).
    Returns on `dst` the concatenation of all shards in rank order (a uint8 tensor), None elsewhere.
    Ranks own contiguous, increasing frame ranges, so concatenation by rank IS the time-ordered merge."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = events_u8.device
    cnt = torch.tensor([n_events], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    pad = torch.zeros(mx * EVENT_BYTES, dtype=torch.uint8, device=dev)
    pad[:n_events * EVENT_BYTES] = events_u8[:n_events * EVENT_BYTES]
    if rank == dst:
        bufs = [torch.empty(mx * EVENT_BYTES, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.gather(pad, bufs, dst=dst, group=group)
        return torch.cat([b[:c * EVENT_BYTES] for b, c in zip(bufs, counts)]), counts
    dist.gather(pad, None, dst=dst, group=group)
    return None, counts


def gather_row_shards(rows, group=None, dst=0):
    """rows: (n, ...) tensor whose leading length differs per rank (device for NCCL, CPU for gloo).  Returns on `dst`
    the concatenation over ranks, in rank order, None elsewhere -- the event-frame sums of a sharded clip
    (the preview's percentile is global over the clip, v2ce.py:262-264, so rank 0 needs all of them)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = rows.device
    cnt = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    pad = torch.zeros((mx,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=dev)
    pad[:rows.shape[0]] = rows
    if rank == dst:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.gather(pad, bufs, dst=dst, group=group)
        return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    dist.gather(pad, None, dst=dst, group=group)
    return None


def shard_schedule(starts, mode, b0, b1, batch_size, seq_len=16):
    """Batches [b0, b1) of a clip's window schedule (v2ce.window_schedule) as a clip of their own:
    (first frame, frame count, is_tail, index of the first frame pair, (window starts relative to `first`, mode)).
    Only the clip's last window is pulled back and trimmed to its last `mode` pairs (v2ce.py:153-154,227-236),
    wherever the window before it lives; every window before the shard emits seq_len pairs."""
    w0, w1 = b0 * batch_size, min(b1 * batch_size, len(starts))
    if w1 <= w0:
        return 0, 0, False, w0 * seq_len, (np.zeros(0, dtype=np.int64), 0)
    first = int(starts[w0])
    frame_count = int(starts[w1 - 1]) + seq_len + 1 - first
    is_tail = (w1 == len(starts))
    return first, frame_count, is_tail, w0 * seq_len, (np.asarray(starts[w0:w1]) - first, mode if is_tail else 0)


def _preview_plane(frames_reader, kw):
    """(H, W) of the voxel planes stream_clip produces for this reader and these settings."""
    probe = np.asarray(frames_reader.read_frames_at_indices([0]))
    height, width = kw.get('height', 260), kw.get('width', 346)
    if kw.get('infer_type', 'center') == 'center':
        return height, width
    return height, int(probe.shape[-1] / probe.shape[-2] * height)       # pano: the resized frame's full width


def stream_clip_sharded(model, frames_reader, frame_count, world, rank, seq_len=16, batch_size=1, to_host=True,
                        preview=None, **kw):
    """Run v2ce.stream_clip on this rank's contiguous share of the batches of a clip and gather the
    event shards on rank 0 (the shards never visit the host on their way).  `frames_reader` duck-types VideoReader.
    Returns (event_stream | None, total events); with to_host=False rank 0 gets the merged stream as a device uint8
    tensor instead of a host recarray.

    preview: a dict to receive the event-frame preview of the WHOLE clip on rank 0 (v2ce.py:241-280) --
    ``preview['frames']`` uint8 (N,H,W,3) BGR and ``preview['upper_bound']``.  Every rank keeps the per-pair sums of its
    shard on the device, rank 0 gathers them (0.72 MB per pair at 346x260), takes the clip-global percentile and
    normalises: identical to the single-process preview.  The ceil / percentile / polarity settings are read from
    **kw like stream_clip's."""
    from . import v2ce as drv
    starts, mode = drv.window_schedule(frame_count, seq_len)
    n_batches = -(-len(starts) // batch_size)
    b0, b1 = shard_range(n_batches, world, rank)
    # pano: every batch is one model call per 346-px width tile (v2ce.py:103-111); the tiles of a window stay on
    # the rank that owns the window, so LDATI sees full-width frames and no voxel exchange is needed
    infer_type = kw.get('infer_type', 'center')
    tiles = 1
    if infer_type == 'pano':
        probe = np.asarray(frames_reader.read_frames_at_indices([0]))
        h0, w0 = probe.shape[-2], probe.shape[-1]
        height = kw.get('height', 260)
        tiles = len(drv.pano_tiles(int(w0 / h0 * height), kw.get('width', 346)))     # width after image_pre_processing
    model.sn_advance(model_calls_before(b0, infer_type, tiles) - model.call_count())

    class _Shard:
        """Presents windows [b0*bs, b1*bs) of the clip as a clip of its own."""

        def __init__(self):
            (self.first, self.frame_count, self.is_tail, self.pair_base, self.schedule) = shard_schedule(
                starts, mode, b0, b1, batch_size, seq_len)

        def read_frames_at_indices(self, idxs):
            return frames_reader.read_frames_at_indices([self.first + i for i in idxs])

    shard = _Shard()
    dev = kw.pop('device', torch.device('cuda', torch.cuda.current_device()))
    sums = None
    if shard.frame_count > 1:
        res = drv.stream_clip(model, vidcap=shard, seq_len=seq_len, batch_size=batch_size,
                              pair_base=shard.pair_base, device=dev, write_event_frames=False, schedule=shard.schedule,
                              events_to_host=False, keep_event_frame_sums=preview is not None, **kw)
        n = res.n_events
        ev = res.event_stream_dev if n else torch.zeros(EVENT_BYTES, dtype=torch.uint8, device=dev)
        sums = res.ef_sums_dev
    else:
        ev, n = torch.zeros(EVENT_BYTES, dtype=torch.uint8, device=dev), 0
    out, counts = gather_event_shards(ev, n)
    if preview is not None:
        from . import event_frames as _ef
        keep = kw.get('keep_polarity', True)
        if sums is None:                       # a rank without windows still takes part in the collective
            h, w = _preview_plane(frames_reader, kw)
            sums = torch.zeros((0, 2 if keep else 1, h, w), dtype=torch.float32, device=dev)
        all_sums = gather_row_shards(sums.contiguous())
        if all_sums is not None:
            with torch.cuda.device(dev):
                ub = _ef.upper_bound(all_sums, kw.get('upper_bound_percentile', 98), kw.get('ceil', 10), keep)
                preview['frames'] = _ef.normalize(all_sums, ub, keep).cpu().numpy()
                preview['upper_bound'] = ub
    if out is not None:
        from .ldati import EVENT_DTYPE
        from .sink import to_host as _sink
        return (_sink(out).view(EVENT_DTYPE) if to_host else out), sum(counts)
    return None, sum(counts)
