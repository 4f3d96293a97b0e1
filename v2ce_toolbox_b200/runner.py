"""BatchRunner: the hot path for one batch of windows, with host transfers overlapped.

    image units (pinned host or device) ─H2D→ V2ce3d ─→ voxels ─→ event-frame sums [+ frames]
                                                         └─→ LDATI count ─(counts D2H)→ emit/sort/pack ─D2H→ host events

Three CUDA streams: the caller's current stream computes, one side stream uploads the next inputs and
another moves results to pinned host buffers while the next batch computes.  Output buffers are double-buffered, so ``submit`` may be
called for batch i+1 before ``wait`` is called for batch i.  This is what ``v2ce.stream_clip`` and
``bench.py`` (e2e) use; it composes the same public calls a user would make one by one
(``V2ce3d.__call__``, ``event_frames.*``, ``LdatiEngine.count/emit``).
"""
import numpy as np
import torch

from . import event_frames as _ef
from . import ldati as _ldati


class Ticket:
    __slots__ = ('slot', 'n_pairs', 'total', 'seg_counts', 'done', 'status', 'ub', 'h2d_bytes', 'd2h_bytes', 'hw')


class BatchRunner:
    def __init__(self, model, device, fps=30, ceil=10, percentile=98, keep_polarity=True, seed=0,
                 per_batch_frames=True, slots=2):
        self.model = model
        self.device = torch.device(device)
        self.fps, self.ceil, self.percentile, self.keep = fps, ceil, percentile, keep_polarity
        self.seed = seed
        self.per_batch_frames = per_batch_frames
        self.eng = _ldati.engine_for(self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)      # D2H of results
        self.h2d_stream = torch.cuda.Stream(device=self.device)       # H2D of inputs (PCIe is full duplex)
        self.slots = slots
        self._ev_dev = [None] * slots
        self._ev_host = [None] * slots
        self._fr_dev = [None] * slots
        self._fr_host = [None] * slots
        self._x_dev = [None] * slots
        self._st_host = [torch.empty(4, dtype=torch.int32, pin_memory=True) for _ in range(slots)]
        self._free = [None] * slots            # event: the side stream has finished reading slot buffers
        self._next = 0
        self.launches = 0
        self.sums = []                          # per-batch event-frame sums (kept for a clip-global percentile)

    def _buf(self, lst, slot, nbytes, pinned=False):
        b = lst[slot]
        if b is None or b.numel() < nbytes:
            n = int(nbytes * 1.25) + 256
            b = torch.empty(n, dtype=torch.uint8, pin_memory=True) if pinned else \
                torch.empty(n, dtype=torch.uint8, device=self.device)
            lst[slot] = b
        return b

    def submit(self, units, pair_base, keep_sums=False):
        """units: (b,L,2,H,W) float32, pinned host (copied on the side stream) or already on the device."""
        t = Ticket()
        slot = self._next
        self._next = (self._next + 1) % self.slots
        t.slot = slot
        cur = torch.cuda.current_stream(self.device)
        if self._free[slot] is not None:
            cur.wait_event(self._free[slot])    # the previous user of this slot has been copied out
        t.h2d_bytes = 0
        if not units.is_cuda:
            t.h2d_bytes = units.numel() * units.element_size()
            with torch.cuda.stream(self.h2d_stream):
                x = units.to(self.device, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.h2d_stream)
            cur.wait_event(ready)
            self._x_dev[slot] = x               # keep alive until the forward has consumed it
        else:
            x = units
        y = self.model(x)
        b, L, _, H, W = y.shape
        n = b * L
        t.n_pairs = n
        t.hw = (H, W)
        vox = y.view(n, 2, 10, H, W)
        sums = _ef.accumulate(vox, self.keep)
        frames = None
        t.ub = None
        if keep_sums:
            self.sums.append(sums)
        if self.per_batch_frames:
            t.ub = _ef.upper_bound(sums, self.percentile, self.ceil, self.keep)
            fr = self._buf(self._fr_dev, slot, n * H * W * 3)
            frames = _ef.normalize(sums, t.ub, self.keep, out=fr[:n * H * W * 3].view(n, H, W, 3))
        params = _ldati.make_params(n, H, W, fps=self.fps, seed=self.seed, frame_base=pair_base, device=self.device,
                                    add_frame_offset=True)
        offs = torch.tensor([int((pair_base + i) * 1 / self.fps * 1e6) for i in range(n)], dtype=torch.int64).to(
            self.device, non_blocking=True)
        l0 = self.eng.launches
        seg = self.eng.count(vox, params)
        seg_host = seg.cpu().numpy()            # the one data-dependent sync: the counts size the output
        total = int(seg_host.sum())
        ev = self._buf(self._ev_dev, slot, max(total, 1) * 13)
        _, status = self.eng.emit(vox, params, total, frame_offsets=offs, out=ev)
        self.launches += self.model.last_launches() + (11 if self.per_batch_frames else 1) + (self.eng.launches - l0)
        t.total, t.seg_counts = total, seg_host
        # results leave on the side stream while the next batch computes
        computed = torch.cuda.Event()
        computed.record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(computed)
            evh = self._buf(self._ev_host, slot, max(total, 1) * 13, pinned=True)
            evh[:total * 13].copy_(ev[:total * 13], non_blocking=True)
            t.d2h_bytes = total * 13 + seg_host.size * 8 + 16
            if frames is not None:
                frh = self._buf(self._fr_host, slot, frames.numel(), pinned=True)
                frh[:frames.numel()].copy_(frames.reshape(-1), non_blocking=True)
                t.d2h_bytes += frames.numel() + 32
            t.status = self._st_host[slot]
            t.status.copy_(status, non_blocking=True)   # pinned: a pageable target would block the host here
            t.done = torch.cuda.Event()
            t.done.record(self.copy_stream)
        self._free[slot] = t.done
        return t

    def wait(self, t, copy=True):
        """Block until batch `t` is on the host.  Returns (events recarray view, frames uint8 (n,H,W,3) | None)."""
        t.done.synchronize()
        _ldati.check_status(t.status.numpy())
        ev = self._ev_host[t.slot][:t.total * 13].numpy().view(_ldati.EVENT_DTYPE)
        fr = None
        if self.per_batch_frames:
            n, (H, W) = t.n_pairs, t.hw
            fr = self._fr_host[t.slot][:n * H * W * 3].numpy().reshape(n, H, W, 3)
        if copy:
            ev = ev.copy()
            fr = fr.copy() if fr is not None else None
        return ev, fr
