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


def init_process_group(backend='nccl', device=None, max_ctas=None, **kw):
    """torch.distributed.init_process_group plus the host-side gloo group the count exchanges use.  `max_ctas` (or
    V2CE_NCCL_MAX_CTAS) caps NCCL's CTAs per operation; the default leaves NCCL's own choice: capped at 4 the merge of
    eight 150 MB shards per step on rank 0 ran at ~100 GB/s and the 8-GPU step went from 10.5 to 17.8 ms
    (profiles/bench_r2_8gpu_a.json) -- the shard transfers need the channels more than the network needs those SMs."""
    if backend == 'nccl':
        import os
        opts = dist.ProcessGroupNCCL.Options()
        cap = max_ctas if max_ctas is not None else os.environ.get('V2CE_NCCL_MAX_CTAS')
        if cap:
            opts.config.max_ctas = int(cap)
            opts.config.min_ctas = 1
        dist.init_process_group('nccl', device_id=device, pg_options=opts, **kw)
        # host-side values (event counts are known on the host: the count pass is read back to size the outputs) are
        # exchanged over a gloo group: no NCCL kernel, no device sync.  The conv kernels are persistent CTAs that fill
        # every SM's shared memory, so an NCCL kernel only starts in the gap between two of them -- a host that waits
        # for a count exchange on the device idles the GPU for that long at every step (0.65 ms of a 10.7 ms step at
        # N = 2, profiles/bench_r2_2gpu_a.json).
        _host_group[0] = dist.new_group(backend='gloo')
        return None
    return dist.init_process_group(backend, **kw)


_host_group = [None]


def exchange_counts(n, group=None, device=None):
    """Every rank's count as a list.  Over the host-side gloo group when init_process_group set one up (no device
    work at all); otherwise one all_gather on `device` and ONE host read."""
    world = dist.get_world_size(group)
    if group is None and _host_group[0] is not None:
        every = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(every, torch.tensor([int(n)], dtype=torch.int64), group=_host_group[0])
        return [int(v.item()) for v in every]
    mine = torch.tensor([int(n)], dtype=torch.int64, device=device)
    every = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(every, mine, group=group)
    return [int(v) for v in every.cpu().tolist()]


def _gather_exact(flat, counts, unit, group, dst, out):
    """Rank-ordered concatenation of the ranks' `flat[:counts[r]*unit]` on `dst`: every shard travels at its exact
    length, point to point, straight into its place in the merged buffer (no padding to the longest shard, no staging
    copy, no torch.cat).  On NCCL nothing here blocks the host: `wait()` orders the current stream behind the
    transfer."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if rank != dst:
        if counts[rank]:
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, flat[:counts[rank] * unit], dst, group)]):
                w.wait()
        return None
    total = sum(counts) * unit
    if out is None or out.numel() < total:
        out = torch.empty(max(total, 1), dtype=flat.dtype, device=flat.device)
    ops, off = [], 0
    for r in range(world):
        n = counts[r] * unit
        if r == rank:
            out[off:off + n].copy_(flat[:n])
        elif n:
            ops.append(dist.P2POp(dist.irecv, out[off:off + n], r, group))
        off += n
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out[:total]


def gather_event_shards(events_u8, n_events, group=None, dst=0, out=None, window=None, slot=0):
    """events_u8: 1-D uint8 tensor holding n_events*13 bytes (device for NCCL, CPU for gloo).
    Returns on `dst` the concatenation of all shards in rank order (a uint8 tensor; a view of `out` when that
    preallocated buffer is large enough), None elsewhere, and the per-rank event counts everywhere.
    Ranks own contiguous, increasing frame ranges, so concatenation by rank IS the time-ordered merge.
    window: a PeerWindow -- the shards are then written into its buffer `slot` by the copy engines (no NCCL kernel) and
    the call returns once every rank's copy has landed."""
    counts = exchange_counts(n_events, group, events_u8.device)
    if window is not None:
        window.push(slot, events_u8, counts, EVENT_BYTES)
        window.fence()
        return (window.view(slot, sum(counts) * EVENT_BYTES) if window.rank == window.dst else None), counts
    return _gather_exact(events_u8, counts, EVENT_BYTES, group, dst, out), counts


class _DeviceBytes:
    """`nbytes` of device memory at `ptr` for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {'shape': (int(nbytes),), 'typestr': '|u1', 'data': (int(ptr), False),
                                         'version': 3, 'strides': None}


class PeerWindow:
    """Gather destination on rank `dst` that every rank of the node writes directly (`buffers` buffers of `nbytes`).
    `dst` allocates it (v2ce_peer_window_alloc), the 64-byte CUDA IPC handles travel over the host group, the other
    ranks map it (v2ce_peer_window_open), and push() is one cudaMemcpyAsync per rank -- the copy engines move the shard
    over NVLink into its place in the merged stream.  Why not NCCL here: the conv kernels are persistent CTAs that fill
    every SM's shared memory, so an NCCL send/recv kernel both waits for a gap between two of them and then holds SMs
    the next conv kernel's CTAs need (forward 9.2 -> 11.0 ms inside the 8-GPU step, profiles/bench_r2_8gpu_b.json); a
    DMA copy needs no SM.  Single node only."""

    def __init__(self, nbytes, buffers=1, dst=0, group=None, device=None):
        from . import _lib
        import ctypes
        self._lib, self._ct = _lib, ctypes
        lib = _lib.load()
        self.world, self.rank, self.dst = dist.get_world_size(group), dist.get_rank(group), dst
        self.group = _host_group[0] if (group is None and _host_group[0] is not None) else group
        self.nbytes = int(nbytes)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.ptrs, self._views = [], []
        err = None
        with torch.cuda.device(self.device):
            for _ in range(buffers):
                box, p = [None], ctypes.c_void_p()
                if self.rank == dst and err is None:
                    h = (ctypes.c_uint8 * 64)()
                    if lib.v2ce_peer_window_alloc(self.nbytes, ctypes.byref(p), h) == 0:
                        box[0] = bytes(h)
                        self._views.append(torch.as_tensor(_DeviceBytes(p.value, self.nbytes), device=self.device))
                    else:
                        err = lib.v2ce_last_error().decode()
                dist.broadcast_object_list(box, src=dst, group=self.group)
                if self.rank != dst and err is None:
                    if box[0] is None:
                        err = 'the destination rank could not allocate the window'
                    elif lib.v2ce_peer_window_open((ctypes.c_uint8 * 64).from_buffer_copy(box[0]), ctypes.byref(p)) != 0:
                        err = lib.v2ce_last_error().decode()
                if err is None:
                    self.ptrs.append(int(p.value))
        # every rank learns whether every rank has the window: a one-sided failure must not leave the others waiting
        every = [None] * self.world
        dist.all_gather_object(every, err, group=self.group)
        if any(e is not None for e in every):
            self._release()
            raise _lib.V2ceError('peer window unavailable: ' + '; '.join(f'rank {r}: {e}' for r, e in enumerate(every) if e))

    def push(self, slot, flat, counts, unit=1):
        """Enqueue, on the current stream, the copy of this rank's `flat[:counts[rank]*unit]` to its place in buffer
        `slot` (rank-ordered, exact lengths).  Nothing waits; fence() (or any later host barrier behind a stream
        synchronize on every rank) makes the merged buffer complete on `dst`."""
        total = sum(counts) * unit
        if total > self.nbytes:
            raise ValueError(f'peer window too small: {total} > {self.nbytes} bytes')
        off, n = sum(counts[:self.rank]) * unit, counts[self.rank] * unit
        if n:
            flat = self._lib.require_cuda(flat, 'shard')
            self._lib.check(self._lib.load().v2ce_peer_copy_async(
                self._ct.c_void_p(self.ptrs[slot] + off), self._ct.c_void_p(flat.data_ptr()), n,
                self._lib.stream_ptr()))

    def fence(self):
        torch.cuda.current_stream(self.device).synchronize()
        dist.barrier(group=self.group)

    def view(self, slot, nbytes=None):
        """dst only: the first `nbytes` of buffer `slot` as a uint8 device tensor."""
        v = self._views[slot]
        return v if nbytes is None else v[:nbytes]

    def _release(self):
        lib = self._lib.load()
        self._views = []
        for p in self.ptrs:
            (lib.v2ce_peer_window_free if self.rank == self.dst else lib.v2ce_peer_window_close)(self._ct.c_void_p(p))
        self.ptrs = []

    def close(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)               # nobody is still copying into a buffer that is about to go
        self._release()


def gather_row_shards(rows, group=None, dst=0):
    """rows: (n, ...) tensor whose leading length differs per rank (device for NCCL, CPU for gloo).  Returns on `dst`
    the concatenation over ranks, in rank order, None elsewhere -- the event-frame sums of a sharded clip
    (the preview's percentile is global over the clip, v2ce.py:262-264, so rank 0 needs all of them)."""
    counts = exchange_counts(rows.shape[0], group, rows.device)
    unit = 1
    for d in rows.shape[1:]:
        unit *= int(d)
    merged = _gather_exact(rows.contiguous().reshape(-1), counts, unit, group, dst, None)
    if merged is None:
        return None
    return merged.reshape((sum(counts),) + tuple(rows.shape[1:]))


_merge_seq = [0]


def merge_event_shards_to_host(events_u8, n_events, group=None, dst=0, workers=8, via=None):
    """The merged, time-ordered event stream of a sharded clip as ONE host array on `dst`.  Returns (uint8 numpy array
    on dst | None, per-rank counts).  Two routes (`via`, default from V2CE_SHARD_MERGE):
      'nccl'  the shards are merged on dst's GPU over NVLink (exact-length point-to-point transfers) and copied down
              once (sink.to_host: pinned double buffering, parallel first touch, huge-page hint);
      'shm'   rank `dst` creates a POSIX shared-memory array of the merged size and every rank copies its own device
              shard into its slice of it -- all PCIe links and all ranks' host threads in parallel, nothing funnels
              through one GPU.  Single node only; falls back to 'nccl' when /dev/shm cannot hold the stream.
    Ranks own contiguous, increasing frame ranges, so the rank-ordered layout IS the merge."""
    import os
    from .sink import to_host as _sink
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = events_u8.device
    counts = exchange_counts(n_events, group, dev)
    total = sum(counts) * EVENT_BYTES
    if world == 1:
        return _sink(events_u8[:total], workers=workers), counts
    # 'nccl' (default): device-side merge on dst over NVLink, then ONE pipelined D2H from there -- measured faster than
    # 'shm' at N = 2 (1.78 s against 3.9 s for the 21 GB of a 9000-frame clip, profiles/bench_r2_2gpu_b.json: first-touch
    # of fresh tmpfs pages costs more than rank 0's single PCIe link saves)
    via = via or os.environ.get('V2CE_SHARD_MERGE', 'nccl')
    if via == 'nccl':
        merged = _gather_exact(events_u8, counts, EVENT_BYTES, group, dst, None)
        if merged is None:
            return None, counts
        return _sink(merged, workers=max(workers, min(32, (os.cpu_count() or 8) // 2))), counts
    _merge_seq[0] += 1
    path = f"/dev/shm/v2ce_merge_{os.environ.get('MASTER_PORT', '0')}_{_merge_seq[0]}"
    ok = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == dst:
        try:
            st = os.statvfs('/dev/shm')
            if st.f_bavail * st.f_frsize > total + (64 << 20) and int(os.environ.get('LOCAL_WORLD_SIZE', world)) == world:
                with open(path, 'wb') as f:
                    f.truncate(max(total, 1))
                ok[0] = 1
        except OSError:
            ok[0] = 0
    dist.broadcast(ok, src=dst, group=group)        # also orders "file exists" before the other ranks open it
    if int(ok.item()) == 0:
        merged, _ = gather_event_shards(events_u8, n_events, group, dst)
        return (_sink(merged, workers=workers) if merged is not None else None), counts
    try:
        off = sum(counts[:rank]) * EVENT_BYTES
        n = counts[rank] * EVENT_BYTES
        if n:
            # every rank allocates its own range of the tmpfs file in the kernel (no page fault per 4 KB from the copy
            # threads: faulting 21 GB of fresh tmpfs pages in from user space ran at 3-4 GB/s per rank)
            fd = os.open(path, os.O_RDWR)
            try:
                os.posix_fallocate(fd, off, n)
            finally:
                os.close(fd)
        mm = np.memmap(path, dtype=np.uint8, mode='r+', shape=(max(total, 1),))
        if n:
            _sink(events_u8[:n], out=mm[off:off + n], workers=workers)
        dist.barrier(group=group)
    finally:
        if rank == dst and os.path.exists(path):
            os.unlink(path)                          # the mapping keeps the pages alive
    return (mm[:total] if rank == dst else None), counts


def close_host_group():
    if _host_group[0] is not None:
        dist.destroy_process_group(_host_group[0])
        _host_group[0] = None


class SharedHostRing:
    """Per-step host destination of a multi-GPU run: `slots` shared-memory arrays that every rank maps and page-locks
    (cudaHostRegister), so each rank's D2H lands directly in ITS slice of the merged, rank-ordered stream of the step
    -- no NCCL gather through one GPU, no second host copy.  place(slot, nbytes) exchanges the ranks' byte counts (one
    gloo all_gather on the host: the counts are host values already, runner._stage_b) and returns this rank's pinned
    slice; layout(slot) gives every rank's (offset, nbytes) of the last placement.  Single node only."""

    def __init__(self, slots, bytes_per_slot, group=None, register=True):
        import os
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.cpu_group = _host_group[0] if _host_group[0] is not None else dist.new_group(backend='gloo')
        self.cap = int(bytes_per_slot)
        self.path = f"/dev/shm/v2ce_ring_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
        total = slots * self.cap
        flag = torch.zeros(1, dtype=torch.int64)
        if self.rank == 0:
            try:
                st = os.statvfs('/dev/shm')
                if st.f_bavail * st.f_frsize > total + (64 << 20):
                    with open(self.path, 'wb') as f:
                        f.truncate(total)
                    flag[0] = 1
            except OSError:
                pass
        dist.broadcast(flag, src=0, group=self.cpu_group)
        if int(flag.item()) == 0:
            raise OSError('/dev/shm cannot hold the shared host ring')
        self.buf = torch.from_file(self.path, shared=True, size=total, dtype=torch.uint8)
        rc = torch.cuda.cudart().cudaHostRegister(self.buf.data_ptr(), total, 0) if register else 0     # False: CPU tests
        ok = torch.tensor([1 if int(rc) == 0 else 0], dtype=torch.int64)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.cpu_group)
        if self.rank == 0:
            os.unlink(self.path)                     # every rank has it mapped; the pages live until the last unmap
        self.registered = register and int(rc) == 0
        if int(ok.item()) == 0:
            self.close()
            raise OSError(f'cudaHostRegister of the shared host ring failed on some rank (rc {int(rc)} here)')
        self.slots = slots
        self._layout = [None] * slots

    def place(self, slot, nbytes):
        mine = torch.tensor([int(nbytes)], dtype=torch.int64)
        every = [torch.zeros(1, dtype=torch.int64) for _ in range(self.world)]
        dist.all_gather(every, mine, group=self.cpu_group)
        sizes = [int(v.item()) for v in every]
        if sum(sizes) > self.cap:
            raise ValueError(f'shared host ring slot too small: {sum(sizes)} > {self.cap} bytes')
        offs = [sum(sizes[:r]) for r in range(self.world)]
        self._layout[slot] = list(zip(offs, sizes))
        base = slot * self.cap + offs[self.rank]
        return self.buf[base:base + max(int(nbytes), 1)]

    def layout(self, slot):
        return self._layout[slot]

    def merged(self, slot):
        """uint8 view of the merged stream of the last placement in `slot` (valid on every rank once all ranks' copies
        have completed)."""
        offs_sizes = self._layout[slot]
        n = offs_sizes[-1][0] + offs_sizes[-1][1]
        return self.buf[slot * self.cap:slot * self.cap + n]

    def close(self):
        if getattr(self, 'registered', False):
            torch.cuda.cudart().cudaHostUnregister(self.buf.data_ptr())
            self.registered = False


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


def pano_tile_owner(n_batches, world, rank):
    """Tile-level sharding of a pano clip with fewer batches than ranks (north_star's "pano tiles" partition): ranks
    form `n_batches` groups of g = world // n_batches; group b runs batch b, its members take the window's 346-px
    tiles round robin and exchange the voxel tiles, and every member then holds the full-width voxels (LDATI needs
    full-width frames).  Returns (batch index or None for an idle rank, member ranks, is_leader); None when the
    window-level split applies (g < 2)."""
    g = world // max(n_batches, 1)
    if n_batches == 0 or g < 2:
        return None
    b = rank // g
    if b >= n_batches:
        return (None, [], False)
    members = list(range(b * g, (b + 1) * g))
    return (b, members, rank == members[0])


def _pano_tiles_shared(members, rank):
    """The pano_fn of a tile-sharing group (v2ce.stream_clip): each member runs the tiles j with members[j % g] == rank
    at the call index the single-process schedule gives them (spectral-norm replay in between), sends them to the other
    members point to point and stitches the full-width voxels exactly like v2ce._pano_device."""
    from . import v2ce as drv

    def fn(model, image_units, width=346):
        tl = drv.pano_tiles(image_units.shape[-1], width)
        g = len(members)
        call0 = model.call_count()                         # every member enters the batch at the same call index
        parts = [None] * len(tl)
        for j, (a, b, keep) in enumerate(tl):
            if members[j % g] == rank:
                if call0 + j > model.call_count():
                    model.sn_advance(call0 + j - model.call_count())
                parts[j] = model(image_units[..., a:b].float().contiguous())
        if call0 + len(tl) > model.call_count():
            model.sn_advance(call0 + len(tl) - model.call_count())
        B, L, _, H, _ = image_units.shape
        ops = []
        for j in range(len(tl)):
            owner = members[j % g]
            if owner == rank:
                ops += [dist.P2POp(dist.isend, parts[j], r) for r in members if r != rank]
            else:
                parts[j] = torch.empty((B, L, 20, H, width), dtype=torch.float32, device=image_units.device)
                ops.append(dist.P2POp(dist.irecv, parts[j], owner))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        parts = [p[..., -keep:] if keep != width else p for p, (a, b, keep) in zip(parts, tl)]
        return torch.cat(parts, dim=-1) if len(parts) > 1 else parts[0]
    return fn


def stream_clip_sharded(model, frames_reader, frame_count, world, rank, seq_len=16, batch_size=1, to_host=True,
                        preview=None, merge=None, **kw):
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
    # the single-process schedule continues the spectral-norm iteration from clip to clip: every rank enters a clip
    # at the same call index (`base`), replays the calls of the batches before its share and, when the clip is done,
    # the calls of the batches after it, so that all ranks leave at base + (calls of the whole clip)
    # fewer batches than ranks: share the tiles of each window among a group of ranks instead (the voxel tiles are
    # exchanged point to point; every member then runs the cheap post-network stages, the group leader reports them)
    import os
    split = pano_tile_owner(n_batches, world, rank) if (infer_type == 'pano' and tiles > 1 and dist.is_initialized() and
                                                        os.environ.get('V2CE_PANO_TILE_SHARDING', '1') != '0') else None
    pano_fn, reports = None, True
    if split is not None:
        b, members, reports = split
        b0, b1 = (b, b + 1) if b is not None else (n_batches, n_batches)
        if b is not None:
            pano_fn = _pano_tiles_shared(members, rank)
    base = model.call_count()
    model.sn_advance(model_calls_before(b0, infer_type, tiles))

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
                              events_to_host=False, keep_event_frame_sums=preview is not None, pano_fn=pano_fn, **kw)
        n = res.n_events if reports else 0                 # tile sharing: only the group leader reports the batch
        ev = res.event_stream_dev if n else torch.zeros(EVENT_BYTES, dtype=torch.uint8, device=dev)
        sums = res.ef_sums_dev if reports else None
    else:
        ev, n = torch.zeros(EVENT_BYTES, dtype=torch.uint8, device=dev), 0
    model.sn_advance(base + model_calls_before(n_batches, infer_type, tiles) - model.call_count())
    if to_host and ev.is_cuda:
        host, counts = merge_event_shards_to_host(ev, n, via=merge)     # merge: 'shm' (default) | 'nccl'; V2CE_SHARD_MERGE
        out = None
    else:
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
    from .ldati import EVENT_DTYPE
    if to_host and ev.is_cuda:
        return (host.view(EVENT_DTYPE) if host is not None else None), sum(counts)
    if out is not None:
        return (out.numpy().view(EVENT_DTYPE) if to_host else out), sum(counts)
    return None, sum(counts)
