"""LDATI (stage 2) -- CPU restatement of the reference's timestamp inference.

TEST INFRASTRUCTURE (see oracle/__init__.py).  numpy float32/float64 arithmetic,
one rounding per operation, in the operation order of the reference:

  relocate_counts      <- /root/reference/scripts/LDATI.py:80-106  (y_relocate, non-bidirectional)
  relocate_counts_bidirectional <- /root/reference/scripts/LDATI.py:91-94,96-104,107-122 (bidirectional=True)
  single_timestamps    <- /root/reference/scripts/LDATI.py:156-165
  slope_params         <- /root/reference/scripts/LDATI.py:13-51,184-192
  pool_counts          <- /root/reference/scripts/LDATI.py:176-183 (pooling_type 'weighted' / 'avg')
  multi_timestamps     <- /root/reference/scripts/LDATI.py:194-196,209-212
  assemble_frame       <- /root/reference/scripts/LDATI.py:217-245,248-310 (pick_elements / pick_and_sort)
  sample_voxel_statistical_oracle <- /root/reference/scripts/LDATI.py:126-214 as called by v2ce.py:356

Two arithmetic flavours exist because the reference is a sequence of torch ops
whose scalar semantics differ between devices:

  flavor='cpu'   what torch-CPU computes (true divisions by Python scalars,
                 arange evaluated in double, torch's own vectorised float32
                 sqrt, which is not correctly rounded in this build).  This is
                 the flavour the golden fixtures pin, because the reference can
                 only be executed on CPU in the build container.
  flavor='cuda'  what torch-CUDA computes (division by a Python scalar is a
                 multiply by the rounded reciprocal, arange evaluated in float32,
                 IEEE sqrt).  The CUDA kernel implements this flavour; the
                 per-op differences are verified against torch-CUDA on the B200
                 by tests/test_gpu_torch_semantics.py.

Tie order: the reference sorts each (frame, bin) segment with an *unstable*
argsort (LDATI.py:297), so the order of equal timestamps is undefined there.
The canonical order used by the oracle and the CUDA path is (timestamp, g) with
g the position in the reference's pre-sort concatenation (neg singles, neg
multis, pos singles, pos multis; row-major pixels; j ascending).
"""
import math
import warnings

import numpy as np

from . import philox

EVENT_DTYPE = np.dtype([('timestamp', '<i8'), ('x', '<i2'), ('y', '<i2'), ('polarity', 'i1')])
assert EVENT_DTYPE.itemsize == 13

F32 = np.float32


class Consts:
    """Host-computed scalar constants."""

    def __init__(self, fps, t0=0, flavor='cuda', nbins=9):
        assert flavor in ('cpu', 'cuda')
        self.fps = fps
        self.t0 = t0
        self.flavor = flavor
        self.C = nbins
        self.frame_step = 1 / fps                       # LDATI.py:145
        self.vs = 1 / fps / nbins                       # LDATI.py:146 (python double)
        self.vs32 = F32(self.vs)
        self.inv_vs32 = F32(1 / self.vs)                # scalar of `1 / voxel_step - ...` (LDATI.py:188)
        self.vs2 = self.vs ** 2
        self.eps6 = F32(1e-6)
        self.eps8 = F32(1e-8)
        # reciprocal-multiply constants of the torch-CUDA flavour
        self.r_fps64 = 1.0 / fps
        self.r_c64 = 1.0 / nbins
        self.r_fps32 = F32(1.0 / fps)
        self.r_c32 = F32(1.0 / nbins)
        self.r_vs2_32 = F32(1.0 / self.vs2)
        self.r6_32 = F32(1.0 / 6)
        self.binstart32 = bin_starts(fps, nbins, flavor)
        # `arange + t0` is a float32 add (LDATI.py:164,209)
        self.binstart_t0_32 = (self.binstart32 + F32(t0)).astype(F32)


def bin_starts(fps, nbins=9, flavor='cuda'):
    """torch.arange(0, 1/fps, 1/fps/nbins) (LDATI.py:164,209) as float32.

    size = ceil((end-start)/step) in double; CPU evaluates start+i*step in double
    and rounds to float32, CUDA evaluates it in float32."""
    frame_step = 1 / fps
    vs = 1 / fps / nbins
    n = int(math.ceil((frame_step - 0) / vs))
    i = np.arange(n)
    if flavor == 'cpu':
        return (0.0 + i.astype(np.float64) * vs).astype(F32)
    return (F32(0) + i.astype(F32) * F32(vs)).astype(F32)


def relocate_counts(y):
    """y (..., 10, H, W) float32 -> (n int64 (...,9,H,W), tend float32 (...,9,H,W))."""
    y = np.asarray(y, dtype=F32)
    C = y.shape[-3]
    n = np.zeros(y.shape[:-3] + (C - 1,) + y.shape[-2:], dtype=np.int64)
    tend = np.zeros(n.shape, dtype=F32)
    debt = np.zeros(y.shape[:-3] + y.shape[-2:], dtype=F32)
    eps6 = F32(1e-6)
    for c in range(C - 1):
        x = y[..., c, :, :] - debt
        nc = np.ceil(x - eps6)
        debt = (nc - x).astype(F32)
        n[..., c, :, :] = nc.astype(np.int64)
        tend[..., c, :, :] = debt
    with np.errstate(invalid='ignore'):
        last = (y[..., C - 1, :, :] - debt).astype(np.int32)      # `.int()` truncation
    n[..., C - 2, :, :] += last
    return n, tend


def relocate_counts_bidirectional(y):
    """y_relocate(y, bidirectional=True) (LDATI.py:91-94,107-122): bins 0..3 carry a debt from the left as in
    relocate_counts, bins 8,7,6 carry a `bless` from the right (seeded with the tenth voxel bin), bin 5 settles
    both, and bin 4 is never written (stays n = 0, tend = 0).  All arithmetic is float32, one rounding per torch op;
    tend is the value the reference stores into its float64 `tendency` tensor."""
    y = np.asarray(y, dtype=F32)
    C = y.shape[-3]
    n = np.zeros(y.shape[:-3] + (C - 1,) + y.shape[-2:], dtype=np.int64)
    tend = np.zeros(n.shape, dtype=F32)
    debt = np.zeros(y.shape[:-3] + y.shape[-2:], dtype=F32)
    eps6 = F32(1e-6)
    with np.errstate(invalid='ignore'):
        for c in range((C - 1) // 2):                       # LDATI.py:96-104
            x = y[..., c, :, :] - debt
            nc = np.ceil(x - eps6)
            debt = (nc - x).astype(F32)
            n[..., c, :, :] = nc.astype(np.int64)
            tend[..., c, :, :] = debt
        bless = y[..., C - 1, :, :]                         # LDATI.py:108
        for c in range(C - 2, C // 2, -1):                  # LDATI.py:109-118
            ys = y[..., c, :, :]
            tend[..., c, :, :] = bless
            fl = np.floor((ys + bless) + eps6)
            bless = ((ys - fl) + bless).astype(F32)
            bless = np.maximum(bless, F32(0))               # torch.clamp(min=0); NaN propagates in both
            n[..., c, :, :] = fl.astype(np.int64)
        c = C // 2                                          # LDATI.py:120-122
        tend[..., c, :, :] = bless - debt
        n[..., c, :, :] = np.ceil((y[..., c, :, :] + bless) - debt).astype(np.int64)
    return n, tend


def single_timestamps(tend, k: Consts):
    """Timestamp (us, int64) of the only event of a pixel-bin with n == 1."""
    t = tend.astype(np.float64)
    if k.flavor == 'cpu':
        t = t / k.fps / k.C
    else:
        t = t * k.r_fps64 * k.r_c64
    shape = [1] * t.ndim
    shape[-3] = k.C
    t = t + k.binstart_t0_32.astype(np.float64).reshape(shape)
    t = t * 1e6
    return np.trunc(t).astype(np.int64)


def pool_counts(n, pooling_type='none', kernel_size=3):
    """y_pooled of LDATI.py:176-183: the (integer) counts as float32, spatially pooled per (frame, polarity, bin) plane.

    'weighted': conv2d with [[1,2,1],[2,4,2],[1,2,1]]/16, zero padding 1 (LDATI.py:177-180) -- dyadic weights on small
                integers, so every accumulation order gives the same float32;
    'avg':      AvgPool2d(kernel_size, stride 1, padding kernel_size//2), count_include_pad -> the zero-padded window sum
                (exact) divided by kernel_size**2 in float32 (LDATI.py:181-182); odd kernel sizes only (an even one
                changes the plane size and the reference's reshape fails)."""
    nf = n.astype(F32)
    if pooling_type == 'none':
        return nf
    H, W = nf.shape[-2:]
    if pooling_type == 'weighted':
        r, w1 = 1, np.array([1, 2, 1], dtype=F32)
        weights = np.outer(w1, w1).astype(F32) / F32(16)
    elif pooling_type == 'avg':
        assert kernel_size % 2 == 1, 'AvgPool2d with an even kernel changes the plane size (the reference fails too)'
        r = kernel_size // 2
        weights = np.ones((kernel_size, kernel_size), dtype=F32)
    else:
        raise ValueError(pooling_type)
    pad = np.zeros(nf.shape[:-2] + (H + 2 * r, W + 2 * r), dtype=F32)
    pad[..., r:r + H, r:r + W] = nf
    out = np.zeros_like(nf)
    for dy in range(2 * r + 1):
        for dx in range(2 * r + 1):
            out = (out + weights[dy, dx] * pad[..., dy:dy + H, dx:dx + W]).astype(F32)
    if pooling_type == 'avg':
        out = (out / F32(kernel_size * kernel_size)).astype(F32)
    return out


def slope_params(n, k: Consts, pooled=None):
    """Per pixel-bin (k, b) of the linear density; n (...,9,H,W) int64.  `pooled`: pool_counts(n, ...) when the
    slope is fitted on spatially pooled counts (LDATI.py:176-186), default the counts themselves."""
    nf = n.astype(F32) if pooled is None else pooled.astype(F32)
    S = np.zeros_like(nf)
    S[..., 1:-1, :, :] = nf[..., 2:, :, :] - nf[..., :-2, :, :]     # reflect pad => 0 at both ends
    num = F32(3) * S - F32(0)
    if k.flavor == 'cpu':
        kraw = num / F32(6)
        kk = kraw / F32(k.vs2)
    else:
        kraw = num * k.r6_32
        kk = kraw * k.r_vs2_32
    with np.errstate(divide='ignore', invalid='ignore'):
        kk = (kk / (nf + k.eps8)).astype(F32)
    b = (k.inv_vs32 - (k.vs32 * kk) * F32(0.5)).astype(F32)
    return kk, b


def _sqrt32(x, flavor):
    if flavor == 'cpu':
        import torch  # torch-CPU's float32 sqrt is not correctly rounded
        return torch.sqrt(torch.from_numpy(np.ascontiguousarray(x))).numpy()
    with np.errstate(invalid='ignore'):
        return np.sqrt(x)


def multi_timestamps(kk, b, u, binstart_t0_32, k: Consts, strategy='slope'):
    """Timestamps for events of pixel-bins with n >= 2 (flat arrays): inverse CDF of the linear density
    ('slope', LDATI.py:194-196) or the raw uniform draw taken as SECONDS ('random', LDATI.py:173-174: the
    reference does not scale it by the bin width, so these events spread over [bin start, bin start + 1 s))."""
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        if strategy == 'random':
            t = u.astype(F32)
        else:
            disc = (b * b + (F32(2) * kk) * u).astype(F32)
            t = ((-b) + _sqrt32(disc, k.flavor)) / kk
            if k.flavor == 'cpu':
                t0 = (u / F32(k.fps)) / F32(k.C)
            else:
                t0 = (u * k.r_fps32) * k.r_c32
            t = np.where(kk == 0, t0, t).astype(F32)
        t = t + binstart_t0_32
        t = t * F32(1e6)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ts = np.trunc(t)
            out = ts.astype(np.int64)
            # NaN float -> int64 is 0x8000000000000000 on x86 (cvttss2si) and on torch-CUDA
            # (measured on B200: tests/test_gpu_torch_semantics.py::test_nan_to_long_on_cuda)
            out[np.isnan(ts)] = np.iinfo(np.int64).min
    return out


def assemble_frame(n_f, ts1_f, kk_f, b_f, frame_index, H, W, k: Consts, draw_fn, strategy='slope'):
    """Events of one frame pair in canonical order.  n_f etc.: (2, 9, H, W)."""
    hw = H * W
    segs = []
    seg_counts = np.zeros(k.C, dtype=np.int64)
    for c in range(k.C):
        parts_ts, parts_pix, parts_pol = [], [], []
        for p in (1, 0):                                    # negative plane first (LDATI.py:290-293)
            nn = n_f[p, c].reshape(hw)
            single = np.nonzero(nn == 1)[0]
            parts_ts.append(ts1_f[p, c].reshape(hw)[single])
            parts_pix.append(single)
            parts_pol.append(np.full(single.shape, 1 - p, dtype=np.int8))
            # additional_events_strategy='none' drops every multi-event pixel-bin (LDATI.py:241-244)
            multi = np.nonzero(nn >= 2)[0] if strategy != 'none' else np.zeros(0, dtype=np.int64)
            cnt = nn[multi]
            pix = np.repeat(multi, cnt)
            start = np.cumsum(cnt) - cnt
            j = np.arange(pix.size, dtype=np.int64) - np.repeat(start, cnt)
            idx = philox.event_index(frame_index, p, c, pix, hw)
            u = draw_fn(idx, j, frame_index, p, c, pix)
            ts_m = multi_timestamps(kk_f[p, c].reshape(hw)[pix], b_f[p, c].reshape(hw)[pix],
                                    u, k.binstart_t0_32[c], k, strategy)
            parts_ts.append(ts_m)
            parts_pix.append(pix)
            parts_pol.append(np.full(pix.shape, 1 - p, dtype=np.int8))
        ts = np.concatenate(parts_ts)
        pix = np.concatenate(parts_pix)
        pol = np.concatenate(parts_pol)
        order = np.argsort(ts, kind='stable')               # canonical (ts, g)
        rec = np.empty(ts.size, dtype=EVENT_DTYPE)
        rec['timestamp'] = ts[order]
        rec['x'] = (pix[order] % W).astype(np.int16)
        rec['y'] = (pix[order] // W).astype(np.int16)
        rec['polarity'] = pol[order]
        segs.append(rec)
        seg_counts[c] = ts.size
    return np.concatenate(segs).view(np.recarray), seg_counts


def sample_voxel_statistical_oracle(y, t0=0, fps=30, seed=0, frame_base=0, flavor='cuda',
                                    draws=None, return_seg_counts=False, additional_events_strategy='slope',
                                    bidirectional=False, pooling_type='none', pooling_kernel_size=3):
    """Oracle of sample_voxel_statistical(y, fps=fps, bidirectional=False | True,
    additional_events_strategy='slope' | 'random' | 'none', pooling_type='none').

    y: (B,2,10,H,W) any real dtype.  draws: optional dense (B,2,9,H,W,M) float32
    array of injected uniforms; default = the Philox stream of oracle/philox.py
    keyed by `seed`, with global frame index frame_base + b."""
    y = np.asarray(y)
    B, P, C, H, W = y.shape
    assert P == 2 and additional_events_strategy in ('slope', 'random', 'none')
    k = Consts(fps, t0, flavor, C - 1)
    n, tend = (relocate_counts_bidirectional if bidirectional else relocate_counts)(y.astype(F32))
    ts1 = single_timestamps(tend, k)
    kk, b = slope_params(n, k, None if pooling_type == 'none' else pool_counts(n, pooling_type, pooling_kernel_size))
    out, counts = [], []
    for f in range(B):
        if draws is None:
            def draw_fn(idx, j, frame, p, c, pix):
                return philox.uniform_from_index(idx, j, seed)
        else:
            def draw_fn(idx, j, frame, p, c, pix, _f=f):
                return draws[_f, p, c].reshape(H * W, -1)[pix, j].astype(F32)
        rec, sc = assemble_frame(n[f], ts1[f], kk[f], b[f], frame_base + f, H, W, k, draw_fn,
                                 strategy=additional_events_strategy)
        out.append(rec)
        counts.append(sc)
    if return_seg_counts:
        return out, np.stack(counts)
    return out


def canonicalize(rec):
    """Re-order a reference result (whose tie order is implementation-defined,
    LDATI.py:297) into a tie-independent form for comparison: within runs of equal
    timestamp *inside a bin segment* events are sorted by (polarity, y, x).

    Segments are concatenated in bin order and timestamps are non-decreasing inside
    a segment, so sorting each maximal run of equal timestamps is segment-safe as
    long as two adjacent segments do not share a timestamp value at their boundary;
    the caller compares timestamp sequences first, which catches that case."""
    rec = np.asarray(rec)
    ts = rec['timestamp']
    if ts.size == 0:
        return rec.copy()
    run = np.concatenate([[0], np.cumsum(ts[1:] != ts[:-1])])
    key = np.lexsort((rec['x'], rec['y'], rec['polarity'], run))
    return rec[key]


def events_equal_modulo_ties(a, b):
    """True when a and b hold the same timestamp sequence and, per run of equal
    timestamps, the same multiset of (x, y, polarity)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return False
    if not np.array_equal(a['timestamp'], b['timestamp']):
        return False
    ca, cb = canonicalize(a), canonicalize(b)
    return all(np.array_equal(ca[f], cb[f]) for f in ('timestamp', 'x', 'y', 'polarity'))
