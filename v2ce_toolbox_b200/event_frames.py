"""Host side of the event-frame preview (v2ce.py:241-280): accumulate -> exact percentile ->
normalise, all on the device; only the cv2 mp4 encoder stays on the host."""
import ctypes
import logging
import math

import numpy as np
import torch

from . import _lib
from ._lib import V2ceError, check, ptr, require_cuda, stream_ptr

logger = logging.getLogger('V2CE')


def accumulate(voxels, keep_polarity=True, out=None):
    """voxels (N,2,10,H,W) float32 CUDA -> sums (N,2,H,W) or (N,1,H,W) float32 CUDA (v2ce.py:255,259)."""
    lib = _lib.load()
    require_cuda(voxels, 'voxels')
    N, P, C, H, W = voxels.shape
    if P != 2 or C != 10:
        raise V2ceError(f'expected (N,2,10,H,W), got {tuple(voxels.shape)}')
    v = voxels.float().contiguous()
    if out is None:
        out = torch.empty((N, 2 if keep_polarity else 1, H, W), dtype=torch.float32, device=v.device)
    check(lib.v2ce_ef_accumulate(ptr(v), N, H, W, 1 if keep_polarity else 0, ptr(out), stream_ptr()))
    return out


def numpy_lerp(a, b, t):
    """numpy's percentile interpolation (lib/_function_base_impl._lerp) in float64."""
    a, b = np.float64(a), np.float64(b)
    d = b - a
    return float(a + d * t if t < 0.5 else b - d * (1 - t))


def percentile_from_order_statistics(n_positive, multiplicity, percentile, value_lo, value_hi):
    """np.percentile(v, percentile) for the array v = the positive sums, each repeated `multiplicity` times
    (np.repeat(efs, 3, axis=1) in gray mode, v2ce.py:260), given the two order statistics the device selected:
    value_lo = v_sorted[floor(vi)], value_hi = v_sorted[min(floor(vi) + 1, len - 1)] with vi = (len - 1) * percentile / 100.
    Host part of v2ce.py:262-264: numpy's float64 linear interpolation."""
    vi = (n_positive * multiplicity - 1) * (percentile / 100.0)
    return numpy_lerp(value_lo, value_hi, vi - math.floor(vi))


def upper_bound(sums, percentile=98, ceil=10, keep_polarity=True):
    """min(np.percentile(sums[sums>0], percentile), ceil) (v2ce.py:262-264) by radix select on the device."""
    lib = _lib.load()
    n = ctypes.c_size_t()
    check(lib.v2ce_ef_select_workspace_bytes(ctypes.byref(n)))
    ws = torch.empty(n.value, dtype=torch.uint8, device=sums.device)
    res = torch.empty(4, dtype=torch.int64, device=sums.device)
    mult = 1 if keep_polarity else 3                 # np.repeat(efs, 3, axis=1), v2ce.py:260
    check(lib.v2ce_ef_select(ptr(sums), sums.numel(), float(percentile), mult, ptr(ws), ws.numel(), ptr(res),
                             stream_ptr()))
    npos, lo, bits_lo, bits_hi = (int(x) for x in res.cpu().numpy())
    if npos == 0:
        raise ValueError('event frames hold no positive value: np.percentile of an empty array')
    a = np.array([bits_lo], dtype=np.uint32).view(np.float32)[0]
    b = np.array([bits_hi], dtype=np.uint32).view(np.float32)[0]
    return min(percentile_from_order_statistics(npos, mult, percentile, a, b), ceil)


def normalize(sums, ub, keep_polarity=True, out=None):
    """sums -> (N,H,W,3) uint8 BGR frames on the device (v2ce.py:267-277)."""
    lib = _lib.load()
    N, _, H, W = sums.shape
    if out is None:
        out = torch.empty((N, H, W, 3), dtype=torch.uint8, device=sums.device)
    check(lib.v2ce_ef_normalize(ptr(sums), N, H, W, 1 if keep_polarity else 0, float(ub), ptr(out), stream_ptr()))
    return out


def event_frames(voxels, ceil=10, upper_bound_percentile=98, keep_polarity=True):
    """Device pipeline; returns (frames uint8 (N,H,W,3) CUDA, upper bound)."""
    sums = accumulate(voxels, keep_polarity)
    ub = upper_bound(sums, upper_bound_percentile, ceil, keep_polarity)
    return normalize(sums, ub, keep_polarity), ub


def write_event_frame_video(voxel_grid, ef_video_path, fps, ceil, upper_bound_percentile=98, keep_polarity=True):
    """Drop-in for v2ce.py:241-280.  voxel_grid: numpy (N,2,10,H,W) (uploaded) or a CUDA tensor."""
    import cv2
    logger.info('Writing event frame video...')
    if isinstance(voxel_grid, np.ndarray):
        voxel_grid = torch.from_numpy(np.ascontiguousarray(voxel_grid, dtype=np.float32)).cuda()
    frames, ub = event_frames(voxel_grid, ceil, upper_bound_percentile, keep_polarity)
    logger.info(f'Upper bound of the event frame value during video writing: {ub}')
    frames = frames.cpu().numpy()
    H, W = frames.shape[1:3]
    video = cv2.VideoWriter(ef_video_path, cv2.VideoWriter_fourcc(*'mp4v'), fps, (W, H))
    for f in frames:
        video.write(f)
    video.release()
    logger.info(f'Event frame video written to {ef_video_path}')
    return frames
