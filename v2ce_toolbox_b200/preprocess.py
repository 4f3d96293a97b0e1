"""Device side of image_pre_processing (/root/reference/v2ce.py:45-64) for frames that need resizing.

    image_units_device(frames_u8 (b, L+1, Hs, Ws) uint8 CUDA, height) -> (b, L, 2, height, int(Ws/Hs*height)) float32

One launch of ``v2ce_image_units`` (csrc/preproc.cu): /255, cv2-exact bilinear resize, pair stacking and
Normalize(0.153, 0.165), bit-identical to the host path (``v2ce.image_pre_processing``, which calls cv2).  Frames that
already have the model's height never come here: ``V2ce3d.forward_frames`` takes them as they are.
"""
import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr


def resized_width(src_h, src_w, height):
    return int(src_w / src_h * height)          # v2ce.py:59


def image_units_device(frames_u8, height=260, out=None):
    require_cuda(frames_u8, 'frames_u8')
    if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4:
        raise _lib.V2ceError(f'expected uint8 frames (b, L+1, H, W), got {frames_u8.dtype} {tuple(frames_u8.shape)}')
    b, L1, sh, sw = frames_u8.shape
    dw = resized_width(sh, sw, height)
    frames_u8 = frames_u8.contiguous()
    with torch.cuda.device(frames_u8.device):
        if out is None:
            out = torch.empty((b, L1 - 1, 2, height, dw), dtype=torch.float32, device=frames_u8.device)
        check(_lib.load().v2ce_image_units(ptr(frames_u8), b, L1, sh, sw, height, dw, ptr(out), stream_ptr()))
    return out
