"""Drop-in for /root/reference/train/scripts/stage2/sample_methods/random_even_sample.py: the 'random' and 'even'
baseline samplers the paper compares LDATI with, on libv2ce_b200.so (csrc/ldati.cu, base_* kernels).

    sample_voxel_baseline(y, t0=0, fps=30, even=False, random=False) -> List[np.recarray]
        y (B,2,10,H,W) on a CUDA device; every value of the TEN bins yields floor(y) events -- at u * delta ('random') or
        j / (floor(y) + 1) * delta ('even') into its bin -- plus one more with probability frac(y); a frame's events are
        sorted by timestamp (:118-170).  Records as in LDATI: timestamp<i8, x<i2, y<i2, polarity i1, itemsize 13.

Extra keyword-only arguments (not in the reference): ``seed`` / ``frame_base`` select the counter-based Philox streams
that replace torch.rand / torch.bernoulli (oracle/baseline_oracle.py), ``flavor`` the device whose ``torch.arange`` the
bin origins follow ('cuda' default, 'cpu').  Order of equal timestamps (undefined in the reference: np.sort on a
structured array): bins ascending, negative before positive plane, integer-part before fractional-part events, pixels
row-major.  Voxels must be finite and, for 'even', non-negative (a value in (-1, 0) divides by zero in the reference
too: its timestamp is -inf): such events raise V2ceError instead of being written."""
import ctypes
import math
from typing import List

import numpy as np
import torch

from .. import _lib, ldati as _ldati
from .._lib import BaselineParams, V2ceError, check, ptr, require_cuda, stream_ptr

NB = 10
_ws = {}


def make_params(n_frames, height, width, fps=30, t0=0, mode='random', seed=0, frame_base=0, flavor='cuda', device='cuda',
                add_frame_offset=False):
    f32 = np.float32
    p = BaselineParams()
    p.height, p.width, p.n_frames = height, width, n_frames
    p.mode = {'random': 1, 'even': 2}[mode]
    p.frame_base = frame_base
    p.seed = seed & 0xFFFFFFFFFFFFFFFF
    p.delta32 = f32(1 / (fps * NB))
    dev = device if flavor == 'cuda' else 'cpu'
    starts = torch.arange(0, 1 / fps, 1 / fps / NB, device=dev).cpu().numpy().astype(f32)      # random_even_sample.py:145
    if starts.shape[0] != NB:
        raise V2ceError(f'torch.arange(0, 1/{fps}, 1/{fps}/10) has {starts.shape[0]} entries; the reference would fail')
    starts = (starts + f32(t0)).astype(f32)
    for c in range(NB):
        p.binstart_t0_32[c] = starts[c]
    p.key_base_us = int(math.floor(float(starts[0]) * 1e6)) - 2
    p.key_span = int(math.ceil(1e6 / fps)) + 8 + 2 * _ldati.KEY_BIAS
    p.add_frame_offset = 1 if add_frame_offset else 0
    return p


def _buf(key, device, nbytes):
    b = _ws.get((key, str(device)))
    if b is None or b.numel() < nbytes:
        b = _ws[(key, str(device))] = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
    return b


def sample_voxel_baseline(y, t0=0, fps=30, even=False, random=False, *, seed=None, frame_base=0, flavor='cuda') -> List[np.recarray]:
    assert (even or random)
    require_cuda(y, 'y')
    B, P, C, H, W = y.shape
    if P != 2 or C != NB:
        raise V2ceError(f'expected y of shape (B,2,10,H,W), got {tuple(y.shape)}')
    vox = y.float().contiguous()
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    lib = _lib.load()
    with torch.cuda.device(vox.device):
        # the reference tests `random` first (random_even_sample.py:135-141): it wins when both flags are set... for the
        # integer part; `even` then overwrites ts.  Both set is not a supported call.
        if even and random:
            raise V2ceError('sample_voxel_baseline: pass exactly one of even / random')
        p = make_params(B, H, W, fps=fps, t0=t0, mode='even' if even else 'random', seed=seed, frame_base=frame_base,
                        flavor=flavor, device=vox.device)
        n = ctypes.c_size_t()
        check(lib.v2ce_baseline_count_workspace_bytes(ctypes.byref(p), ctypes.byref(n)))
        cws = _buf('count', vox.device, n.value)
        counts = torch.empty(B, dtype=torch.int64, device=vox.device)
        check(lib.v2ce_baseline_count(ptr(vox), ctypes.byref(p), ptr(cws), cws.numel(), ptr(counts), stream_ptr()))
        counts_host = counts.cpu().numpy()
        total = int(counts_host.sum())
        check(lib.v2ce_baseline_emit_workspace_bytes(ctypes.byref(p), total, ctypes.byref(n)))
        ews = _buf('emit', vox.device, n.value)
        out = torch.empty(max(total, 1) * 13, dtype=torch.uint8, device=vox.device)
        status = torch.zeros(4, dtype=torch.int32, device=vox.device)
        check(lib.v2ce_baseline_emit(ptr(vox), ctypes.byref(p), ptr(cws), ptr(ews), ews.numel(), None, total, ptr(out),
                                     ptr(status), stream_ptr()))
        host = out[:total * 13].cpu().numpy()
        st = status.cpu().numpy()
    if int(st[0]) != 0:
        raise V2ceError(f'sample_voxel_baseline: {int(st[0])} timestamps fell outside the frame (non-finite voxels, or '
                        f"negative voxels with even=True: the reference's timestamp is -inf there)")
    rec = host.view(_ldati.EVENT_DTYPE)
    res, start = [], 0
    for k in counts_host:
        res.append(rec[start:start + int(k)].view(np.recarray))
        start += int(k)
    return res
