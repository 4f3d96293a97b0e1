"""Drop-in for /root/reference/scripts/LDATI.py: same public name, arguments, defaults and
return type; the tensor work runs in libv2ce_b200.so (csrc/ldati.cu).

    sample_voxel_statistical(y, t0=0, fps=30, pooling_type='none', pooling_kernel_size=3,
                             additional_events_strategy='slope', bidirectional=False)
        -> List[np.recarray]  (len B; dtype timestamp<i8, x<i2, y<i2, polarity i1; itemsize 13;
                               timestamps relative to the frame start, like LDATI.py:308)

Extra keyword-only arguments (not in the reference): ``seed`` / ``frame_base`` select the
counter-based Philox stream that replaces ``torch.rand`` (LDATI.py:171), ``draws`` injects a
dense (B,2,9,H,W,M) tensor of uniforms instead, ``flavor`` picks the torch-CUDA ('cuda',
default: what the reference computes on a GPU) or torch-CPU ('cpu') scalar semantics.

Options: ``additional_events_strategy`` 'slope' (what v2ce.py:356 uses), 'random' and 'none', and
``bidirectional`` False / True, and ``pooling_type`` 'none' / 'weighted' / 'avg' are all implemented by the same
kernels.  With ``bidirectional=True`` a tenth-bin voxel value above
``ldati.BIDIR_MAX_TENDENCY`` (1024 events in one pixel-bin) raises V2ceError.
"""
import logging
from typing import List

import numpy as np
import torch

from .. import ldati as _ldati
from .._lib import V2ceError, require_cuda

logger = logging.getLogger(__name__)


def sample_voxel_statistical(y, t0=0, fps=30, pooling_type='none', pooling_kernel_size=3,
                             additional_events_strategy='slope', bidirectional=False, *,
                             seed=None, frame_base=0, draws=None, flavor='cuda') -> List[np.recarray]:
    assert pooling_type in ['avg', 'weighted', 'none']
    assert additional_events_strategy in ['none', 'random', 'slope']
    require_cuda(y, 'y')
    B, P, C, H, W = y.shape
    if P != 2 or C != 10:
        raise V2ceError(f'expected y of shape (B,2,10,H,W), got {tuple(y.shape)}')
    vox = y.float().contiguous()
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())       # follows torch.manual_seed like torch.rand does
    if draws is not None:
        draws = require_cuda(draws, 'draws').float().contiguous()
        if tuple(draws.shape[:5]) != (B, 2, 9, H, W):
            raise V2ceError(f'draws must be (B,2,9,H,W,M), got {tuple(draws.shape)}')
    with torch.cuda.device(vox.device):
        eng = _ldati.engine_for(vox.device)
        params = _ldati.make_params(B, H, W, fps=fps, t0=t0, seed=seed, frame_base=frame_base, flavor=flavor,
                                    device=vox.device, additional_events_strategy=additional_events_strategy,
                                    bidirectional=bool(bidirectional), pooling_type=pooling_type,
                                    pooling_kernel_size=pooling_kernel_size)
        events, seg_counts, status = eng.run(vox, params, draws=draws)
        total = int(seg_counts.sum())
        host = torch.empty(total * 13, dtype=torch.uint8, pin_memory=True)
        host.copy_(events[:total * 13], non_blocking=True)
        status_host = status.cpu()                                 # synchronises the stream
    _ldati.check_status(status_host.numpy())
    logger.debug(f'LDATI: {total} events in {B} frames')
    return _ldati.split_frames(host.numpy(), seg_counts)
