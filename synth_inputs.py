"""Deterministic synthetic inputs shared by tests, golden generation, bench.py and tools/.

TEST / BENCH INFRASTRUCTURE (input generators only: no algorithm of the hot path lives here, and the product package
never imports it).  Nothing here comes from the reference; the shapes
and distributions are the ones SURVEY.md section 8d names:

  * ``make_state_dict``  random V2ce3d checkpoint with the reference's key layout
    (218 tensors, 52,916,466 elements) -- the Google-Drive checkpoint is not
    available offline.  ``init='reference'`` mimics ``UNet3D.init_weights``
    (kaiming_normal_ with a=10, BN identity); ``init='lively'`` draws BN statistics,
    biases and larger conv gains so every epilogue term and the multi-event LDATI
    path are exercised.
  * ``make_video``       blurred-noise texture translating 2 px/frame with a gain ramp.
  * ``make_voxels``      the LDATI microbench distributions: 'rand' and 'randint'
    (the two the reference's own __main__ uses, scripts/LDATI.py:332,346) and 'sparse'.
"""
import math

import numpy as np
import torch


def layer_table():
    """(name, cin, cout, k, spectral_norm, has_bias) for the 32 convs, reference key order."""
    rows = [('UNet.head.conv3d', 2, 32, 3, False, True)]
    ch = [32, 64, 128, 256, 512]
    for i in range(4):
        p = f'UNet.encoders.{i}'
        rows += [(p + '.conv1', ch[i], ch[i + 1], 3, False, False),
                 (p + '.conv2', ch[i + 1], ch[i + 1], 3, False, False),
                 (p + '.downsample.0', ch[i], ch[i + 1], 1, False, True)]
    for i in range(2):
        p = f'UNet.resblocks.{i}'
        rows += [(p + '.conv1', 512, 512, 3, True, False),
                 (p + '.conv2', 512, 512, 3, True, False),
                 (p + '.downsample.0', 512, 512, 1, False, True)]
    dec_in = [768, 384, 192, 96]
    dec_out = [256, 128, 64, 32]
    for i in range(4):
        p = f'UNet.decoders.{i}'
        rows += [(p + '.conv1', dec_in[i], dec_out[i], 3, True, False),
                 (p + '.conv2', dec_out[i], dec_out[i], 3, True, False),
                 (p + '.downsample.0', dec_in[i], dec_out[i], 1, False, True)]
    rows.append(('UNet.pred.conv3d', 32, 20, 1, False, True))
    return rows


def make_state_dict(seed=0, init='reference'):
    g = torch.Generator(device='cpu').manual_seed(int(seed))
    sd = {}
    lively = init == 'lively'

    def randn(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32)

    def rand(*shape):
        return torch.rand(*shape, generator=g, dtype=torch.float32)

    def bn(prefix, c):
        if lively:
            sd[prefix + '.weight'] = 0.6 + 0.8 * rand(c)
            sd[prefix + '.bias'] = 0.05 * randn(c)
            sd[prefix + '.running_mean'] = 0.05 * randn(c)
            sd[prefix + '.running_var'] = 0.5 + rand(c)
        else:
            sd[prefix + '.weight'] = torch.ones(c)
            sd[prefix + '.bias'] = torch.zeros(c)
            sd[prefix + '.running_mean'] = torch.zeros(c)
            sd[prefix + '.running_var'] = torch.ones(c)
        sd[prefix + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)

    for name, cin, cout, k, sn, bias in layer_table():
        fan_in = cin * k ** 3
        if lively:
            std = math.sqrt(2.0 / fan_in) * (1.6 if sn else 0.9)
        else:
            std = math.sqrt(2.0 / (1 + 10.0 ** 2)) / math.sqrt(fan_in)      # kaiming_normal_(w, a=10)
        w = std * randn(cout, cin, k, k, k)
        if sn:
            u = randn(cout)
            v = randn(fan_in)
            sd[name + '.module.weight_u'] = u / (u.norm() + 1e-12)
            sd[name + '.module.weight_v'] = v / (v.norm() + 1e-12)
            sd[name + '.module.weight_bar'] = w
        else:
            sd[name + '.weight'] = w
            if bias:
                sd[name + '.bias'] = 0.05 * randn(cout) if lively else torch.zeros(cout)
        if name.endswith('.conv1'):
            bn(name[:-len('.conv1')] + '.bn1', cout)
        elif name.endswith('.conv2'):
            bn(name[:-len('.conv2')] + '.bn2', cout)
        elif name.endswith('.downsample.0'):
            bn(name[:-len('.0')] + '.1', cout)
    if lively:
        sd['UNet.pred.conv3d.bias'] = 0.3 * randn(20) - 0.2
    return sd


def order_like(sd, reference_keys):
    return {k: sd[k] for k in reference_keys}


def _blur1d(a, sigma, axis):
    r = int(3 * sigma + 0.5)
    x = np.arange(-r, r + 1)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    a = np.moveaxis(a, axis, -1)
    pad = np.pad(a, [(0, 0)] * (a.ndim - 1) + [(r, r)], mode='wrap')
    out = np.zeros_like(a)
    for i, w in enumerate(k):
        out += w * pad[..., i:i + a.shape[-1]]
    return np.moveaxis(out, -1, axis)


def make_video(n_frames, height=260, width=346, seed=0, shift=2):
    """uint8 gray frames (n, H, W): blurred noise texture translating `shift` px/frame,
    global gain ramp 0.9 -> 1.1."""
    rng = np.random.default_rng(seed)
    period = 64
    base = rng.random((height, width + period)).astype(np.float64)
    base = _blur1d(_blur1d(base, 3.0, 0), 3.0, 1)
    base = (base - base.min()) / (base.max() - base.min())
    frames = np.empty((n_frames, height, width), dtype=np.uint8)
    for i in range(n_frames):
        off = (i * shift) % period
        gain = 0.9 + 0.2 * (i / max(n_frames - 1, 1))
        frames[i] = np.clip(base[:, off:off + width] * 255.0 * gain, 0, 255).astype(np.uint8)
    return frames


class SynthVideoReader:
    """make_video's clip, frame by frame on demand (duck-types VideoReader like FakeVideoReader): a rank of a sharded
    9000-frame or 1080p clip synthesises only the frames it reads.  ``repeat`` > 1 tiles a small texture up by pixel
    replication (a 1080p source from a 270x480 texture)."""

    def __init__(self, n_frames, height=260, width=346, seed=0, shift=2, repeat=1):
        assert height % repeat == 0 and width % repeat == 0
        rng = np.random.default_rng(seed)
        self.period, self.shift, self.repeat = 64, shift, repeat
        self.h, self.w = height // repeat, width // repeat
        base = rng.random((self.h, self.w + self.period)).astype(np.float64)
        base = _blur1d(_blur1d(base, 3.0, 0), 3.0, 1)
        self.base = (base - base.min()) / (base.max() - base.min())
        self.frame_count = n_frames
        self.shape = (height, width)

    def cache_range(self, a, b):
        """Materialise frames [a, b) so that reading them is a memory copy (benchmarks exclude the synthesis)."""
        a, b = max(int(a), 0), min(int(b), self.frame_count)
        self._cache_first = a
        self._cache = np.stack([self.frame(i) for i in range(a, b)], axis=0) if b > a else None

    _cache, _cache_first = None, 0

    def frame(self, i):
        i = max(int(i), 0)
        if self._cache is not None and self._cache_first <= i < self._cache_first + len(self._cache):
            return self._cache[i - self._cache_first]
        off = (i * self.shift) % self.period
        gain = 0.9 + 0.2 * (i / max(self.frame_count - 1, 1))
        f = np.clip(self.base[:, off:off + self.w] * 255.0 * gain, 0, 255).astype(np.uint8)
        if self.repeat > 1:
            f = np.repeat(np.repeat(f, self.repeat, axis=0), self.repeat, axis=1)
        return f

    def read_frames_at_indices(self, idxs):
        idxs = list(idxs)
        if (self._cache is not None and len(idxs) > 0 and idxs[0] >= self._cache_first and
                idxs[-1] < self._cache_first + len(self._cache) and idxs == list(range(idxs[0], idxs[-1] + 1))):
            return self._cache[idxs[0] - self._cache_first:idxs[-1] + 1 - self._cache_first]     # a view: no copy
        return np.stack([self.frame(i) for i in idxs], axis=0)


def make_voxels(kind, n, height=260, width=346, seed=42):
    """(n,2,10,H,W) float32 voxel grids of the LDATI microbench distributions."""
    rng = np.random.default_rng(seed)
    shape = (n, 2, 10, height, width)
    if kind == 'rand':
        return rng.random(shape, dtype=np.float32)
    if kind == 'randint':
        return rng.integers(0, 10, shape, dtype=np.int16).astype(np.float32)
    if kind == 'sparse':
        return (np.float32(0.015) * rng.random(shape, dtype=np.float32)).astype(np.float32)
    if kind == 'mixed':
        return (np.float32(3) * rng.random(shape, dtype=np.float32) ** 3).astype(np.float32)
    raise ValueError(kind)


class FakeVideoReader:
    """In-memory stand-in for scripts.video_reader.VideoReader: what video_to_voxels / stream_clip need from it
    (v2ce.py:149,170) -- ``frame_count`` and ``read_frames_at_indices`` (negative indices clamp to the first frame,
    like the reference reader's seek)."""

    def __init__(self, frames):
        self.frames = np.asarray(frames)
        self.frame_count = self.frames.shape[0]

    def read_frames_at_indices(self, idxs):
        return np.stack([self.frames[max(i, 0)] for i in idxs], axis=0)
