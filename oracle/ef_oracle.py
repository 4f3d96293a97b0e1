"""Event-frame preview (``--write_event_frame_video``) -- CPU restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
/root/reference/v2ce.py:241-280 (write_event_frame_video):

  accumulate   <- v2ce.py:255 / 259-260   per-pair sum of the 10 bins (fp32, sequential)
  upper_bound  <- v2ce.py:262-264         min(percentile of the positive sums, ceil)
  to_bgr_u8    <- v2ce.py:267-277         clip / normalise (float64) -> uint8, RGB->BGR

The reference's "event frame" is built from the predicted *voxels*, not from the
LDATI events.
"""
import numpy as np


def accumulate(voxel, keep_polarity=True):
    """voxel (N,2,10,H,W) fp32 -> sums (N,2,H,W) fp32, or (N,1,H,W) in gray mode.

    numpy reduces a non-contiguous axis by sequential accumulation, which is what
    the explicit loops below restate (left to right over the bins; gray mode adds
    the polarity planes after their bin sums: (((p0b0+p1b0)+p0b1)+... is NOT what
    numpy does -- it reduces axis 1 then axis 2, see _gray)."""
    v = np.asarray(voxel, dtype=np.float32)
    if keep_polarity:
        s = v[:, :, 0].copy()
        for c in range(1, v.shape[2]):
            s = s + v[:, :, c]
        return s
    return _gray(v)


def _gray(v):
    # np.sum(axis=(1,2)) on a C-contiguous (N,2,10,H,W) array: numpy's multi-axis
    # add.reduce iterates with the outer reduced axis (polarity) slowest, i.e. the
    # accumulator visits p0b0..p0b9 then p1b0..p1b9 in order.
    s = v[:, 0, 0].copy()
    first = True
    for p in range(v.shape[1]):
        for c in range(v.shape[2]):
            if first:
                first = False
                continue
            s = s + v[:, p, c]
    return s[:, None]


def order_statistic_pair(values, q):
    """The two order statistics numpy's linear-interpolated percentile reads, and the weight.
    values: 1-D array of the positive sums; q in [0,100]."""
    n = values.size
    vi = (n - 1) * (q / 100.0)
    lo = int(np.floor(vi))
    hi = min(lo + 1, n - 1)
    t = vi - lo
    part = np.partition(values, [lo, hi])
    return part[lo], part[hi], t, n


def lerp_percentile(a, b, t):
    """numpy's _lerp (lib/_function_base_impl.py): a+(b-a)t for t<0.5 else b-(b-a)(1-t), in float64."""
    a = np.float64(a)
    b = np.float64(b)
    d = b - a
    return a + d * t if t < 0.5 else b - d * (1 - t)


def upper_bound(sums, percentile=98, ceil=10, keep_polarity=True):
    """min(np.percentile(positive sums, percentile), ceil) (v2ce.py:262-264).

    In RGB mode the reference concatenates a zero blue channel (never > 0); in gray
    mode it repeats the single channel three times, which triples every multiplicity."""
    flat = np.asarray(sums, dtype=np.float32).reshape(-1)
    pos = flat[flat > 0]
    if not keep_polarity:
        pos = np.repeat(pos, 3)
    a, b, t, n = order_statistic_pair(pos, percentile)
    return min(lerp_percentile(a, b, t), ceil)


def to_bgr_u8(sums, ub, keep_polarity=True):
    """sums (N,2|1,H,W) fp32 -> frames (N,H,W,3) uint8 in BGR order as handed to cv2.VideoWriter."""
    s = np.asarray(sums, dtype=np.float64)
    x = np.clip(s, 0, ub) / ub
    u8 = (x * 255).astype(np.uint8)                     # truncation
    N, _, H, W = u8.shape
    out = np.zeros((N, H, W, 3), dtype=np.uint8)
    if keep_polarity:
        out[..., 2] = u8[:, 0]                          # R = p-index 0
        out[..., 1] = u8[:, 1]                          # G = p-index 1
    else:
        out[...] = u8[:, 0][..., None]
    return out


def event_frames_oracle(voxel, ceil=10, percentile=98, keep_polarity=True):
    sums = accumulate(voxel, keep_polarity)
    ub = upper_bound(sums, percentile, ceil, keep_polarity)
    return to_bgr_u8(sums, ub, keep_polarity), ub, sums
