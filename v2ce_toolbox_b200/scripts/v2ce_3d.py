"""Drop-in for /root/reference/scripts/v2ce_3d.py: ``V2ce3d`` with the reference's constructor
arguments, state-dict key layout and forward contract

    forward(x: (B, L, 2, H, W) float32 CUDA) -> (B, L, 20, H, W) float32 CUDA, values >= 0

running on libv2ce_b200.so (csrc/unet.cu, csrc/conv_igemm.cuh).  Like the reference, every
forward advances the spectral-norm power iteration by one step (scripts/spectral_norm.py:62-64),
so outputs depend on the model-call index.
"""
import ctypes
import logging

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from .._lib import V2ceError, check, ptr, require_cuda, stream_ptr

logger = logging.getLogger(__name__)

_SKIP_SUFFIX = ('num_batches_tracked',)


class V2ce3d(nn.Module):
    def __init__(self, in_channels=2, out_channels=20):
        super().__init__()
        if in_channels != 2 or out_channels != 20:
            raise V2ceError('the B200 path implements the released V2ce3d(in_channels=2, out_channels=20)')
        self._lib = _lib.load()
        self._handle = None
        self._state = None
        self._device = None
        self._ws = None

    # -- checkpoint -----------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True):
        """Accepts the reference's state_dict (218 entries, keys ``UNet.*``).  Returns torch's ``_IncompatibleKeys``
        result like ``nn.Module.load_state_dict``; missing or mis-sized tensors raise when the device model is built
        (`strict` or not: the network cannot run without them)."""
        from torch.nn.modules.module import _IncompatibleKeys
        self._state = {k: v.detach().to('cpu', torch.float32).contiguous()
                       for k, v in state_dict.items() if not k.endswith(_SKIP_SUFFIX)}
        unexpected = [k for k in self._state if not k.startswith('UNet.')]
        if strict and unexpected:
            raise V2ceError(f'unexpected key(s) in state_dict: {unexpected[:4]}')
        self._release()
        if self._device is not None:
            self._build()
        return _IncompatibleKeys([], unexpected)

    def state_dict(self, *a, **k):
        return dict(self._state or {})

    def to(self, device=None, *a, **k):
        if device is None:
            return self
        device = torch.device(device)
        if device.type != 'cuda':
            raise V2ceError('V2ce3d (B200) has no CPU path')
        self._device = torch.device('cuda', device.index if device.index is not None else torch.cuda.current_device())
        if self._state is not None and self._handle is None:
            self._build()
        return self

    def cuda(self, device=None):
        return self.to('cuda' if device is None else f'cuda:{device}' if isinstance(device, int) else device)

    def _build(self):
        h = ctypes.c_void_p()
        with torch.cuda.device(self._device):
            check(self._lib.v2ce_model_create(ctypes.byref(h), self._device.index))
            for name, t in self._state.items():
                arr = t.numpy()
                shape = (ctypes.c_int64 * max(arr.ndim, 1))(*arr.shape)
                check(self._lib.v2ce_model_set_tensor(h, name.encode(), arr.ctypes.data_as(ctypes.c_void_p), shape,
                                                      arr.ndim))
            check(self._lib.v2ce_model_finalize(h))
        self._handle = h

    def _release(self):
        if self._handle is not None:
            self._lib.v2ce_model_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # -- forward --------------------------------------------------------------------------
    def forward(self, x):
        require_cuda(x, 'x')
        if self._handle is None:
            if self._state is None:
                raise V2ceError('V2ce3d: load_state_dict() before forward()')
            self.to(x.device)
        B, L, C, H, W = x.shape
        if C != 2:
            raise V2ceError(f'expected x of shape (B,L,2,H,W), got {tuple(x.shape)}')
        x = x.float().contiguous()
        with torch.cuda.device(x.device):
            n = ctypes.c_size_t()
            check(self._lib.v2ce_model_workspace_bytes(self._handle, B, L, H, W, ctypes.byref(n)))
            if self._ws is None or self._ws.numel() < n.value:
                self._ws = None
                self._ws = torch.empty(n.value, dtype=torch.uint8, device=x.device)
            y = torch.empty((B, L, 20, H, W), dtype=torch.float32, device=x.device)
            check(self._lib.v2ce_model_forward(self._handle, ptr(x), ptr(y), B, L, H, W, ptr(self._ws),
                                               self._ws.numel(), stream_ptr()))
        return y

    def forward_frames(self, frames):
        """frames (B, L+1, H, W) uint8 CUDA, window b = L+1 consecutive gray frames at the model's resolution ->
        (B, L, 20, H, W) float32.  Equals ``forward(image_pre_processing(frames))`` bit for bit (v2ce.py:45-64);
        the pre-processing runs inside the head conv, so a window costs 1/8 of the H2D bytes of its image units."""
        require_cuda(frames, 'frames')
        if frames.dtype != torch.uint8 or frames.dim() != 4:
            raise V2ceError(f'expected uint8 frames of shape (B,L+1,H,W), got {frames.dtype} {tuple(frames.shape)}')
        if self._handle is None:
            if self._state is None:
                raise V2ceError('V2ce3d: load_state_dict() before forward()')
            self.to(frames.device)
        B, L1, H, W = frames.shape
        L = L1 - 1
        frames = frames.contiguous()
        with torch.cuda.device(frames.device):
            n = ctypes.c_size_t()
            check(self._lib.v2ce_model_workspace_bytes(self._handle, B, L, H, W, ctypes.byref(n)))
            if self._ws is None or self._ws.numel() < n.value:
                self._ws = None
                self._ws = torch.empty(n.value, dtype=torch.uint8, device=frames.device)
            y = torch.empty((B, L, 20, H, W), dtype=torch.float32, device=frames.device)
            check(self._lib.v2ce_model_forward_frames(self._handle, ptr(frames), ptr(y), B, L, H, W, ptr(self._ws),
                                                      self._ws.numel(), stream_ptr()))
        return y

    # -- spectral-norm state (multi-GPU replay, tests) --------------------------------------
    def last_sigmas(self):
        out = (ctypes.c_float * 12)()
        check(self._lib.v2ce_model_last_sigmas(self._handle, out))
        return np.array(out[:], dtype=np.float32)

    def call_count(self):
        n = ctypes.c_int64()
        check(self._lib.v2ce_model_call_count(self._handle, ctypes.byref(n)))
        return n.value

    def sn_advance(self, n_calls):
        """Advance the power iteration as if `n_calls` forwards had run (windows owned by other ranks)."""
        with torch.cuda.device(self._device):
            check(self._lib.v2ce_model_sn_advance(self._handle, int(n_calls), stream_ptr()))

    def last_launches(self):
        n = ctypes.c_int32()
        check(self._lib.v2ce_model_last_launches(self._handle, ctypes.byref(n)))
        return n.value

    def set_option(self, key, value):
        check(self._lib.v2ce_model_set_option(self._handle, key.encode(), int(value)))

    def layer_times(self):
        """[(layer name, ms)] of the last forward (after set_option('layer_timing', 1))."""
        cap = 64
        ms = (ctypes.c_float * cap)()
        names = ctypes.create_string_buffer(cap * 48)
        n = ctypes.c_int32()
        check(self._lib.v2ce_model_layer_times(self._handle, cap, ms, names, ctypes.byref(n)))
        return [(names.raw[i * 48:(i + 1) * 48].split(b'\0', 1)[0].decode(), float(ms[i])) for i in range(n.value)]
