"""Baseline samplers 'random' / 'even' (stage-2 comparison methods) -- CPU restatement.

TEST INFRASTRUCTURE (see oracle/__init__.py).  numpy float32 arithmetic, one rounding per torch op, in the order of

  sample_voxel_baseline   <- /root/reference/train/scripts/stage2/sample_methods/random_even_sample.py:118-170
  pick_elements(_bn) / pick_and_sort <- same file :20-115

What the reference does: every voxel value y[b,p,c] (10 bins here, not 9: no count relocation) is split into
floor(y) events -- timestamps `u * delta` ('random', one uniform draw each) or `j / (floor(y) + 1) * delta` ('even') into
the bin -- plus one more event with probability frac(y) (torch.bernoulli) at `u * delta` / `floor(y) / (floor(y)+1) * delta`;
per frame everything is sorted by timestamp (np.sort(order='timestamp'): the order of equal timestamps is undefined).

Draws (SURVEY.md F7 applied to this sampler): Philox4x32-10 keyed by `seed`, counter (pixel-bin index over 10 bins, j >> 2,
stream) with stream 0 = integer-part uniforms, 1 = fractional-part uniform, 2 = the Bernoulli uniform (selected when
u < frac, which is how torch.bernoulli(p) consumes a uniform).  Canonical order: (timestamp, g), g = position in
[bin c][neg, pos][integer-part events (h, w, j), then fractional-part events (h, w)].

Flavours as in ldati_oracle: 'cpu' = torch-CPU (arange evaluated in double), 'cuda' = torch-CUDA (arange in float32).
"""
import numpy as np

from . import philox
from .ldati_oracle import EVENT_DTYPE, F32, bin_starts

NB = 10


def pixel_bin_index(frame, p, c, pix, hw):
    return ((np.uint64(frame) * np.uint64(2) + np.uint64(p)) * np.uint64(NB) + np.uint64(c)) * np.uint64(hw) + \
        np.asarray(pix, dtype=np.uint64)


def sample_voxel_baseline_oracle(y, t0=0, fps=30, even=False, random=False, seed=0, frame_base=0, flavor='cuda'):
    assert even or random
    y = np.asarray(y).astype(F32)
    B, P, C, H, W = y.shape
    assert P == 2 and C == NB
    hw = H * W
    delta = 1 / (fps * C)                                    # python double; torch multiplies by float32(delta)
    d32 = F32(delta)
    starts = (bin_starts(fps, C, flavor) + F32(t0)).astype(F32)
    out = []
    for b in range(B):
        ts_all, x_all, y_all, p_all = [], [], [], []
        frame = frame_base + b
        for c in range(C):
            for p in (1, 0):                                 # negative plane first -> polarity 0
                v = y[b, p, c].reshape(hw)
                ip = np.floor(v)
                frac = (v - ip).astype(F32)
                pix = np.arange(hw)
                idx = pixel_bin_index(frame, p, c, pix, hw)
                # integer part: j < floor(y)
                n = np.where(ip > 0, ip, 0).astype(np.int64)
                rep = np.repeat(pix, n)
                j = (np.arange(rep.size) - np.repeat(np.cumsum(n) - n, n)).astype(np.int64)
                if random:
                    t = (philox.uniform_from_index(idx[rep], j, seed, 0) * d32).astype(F32)
                else:
                    t = ((j.astype(F32) / (ip[rep] + F32(1))).astype(F32) * d32).astype(F32)
                t = ((t + starts[c]).astype(F32) * F32(1e6)).astype(F32)
                ts_all.append(t.astype(np.int64))
                x_all.append((rep % W).astype(np.int16)); y_all.append((rep // W).astype(np.int16))
                p_all.append(np.full(rep.size, 1 - p, np.int8))
                # fractional part: Bernoulli(frac)
                sel = philox.uniform_from_index(idx, np.zeros(hw, np.int64), seed, 2) < frac
                sp = pix[sel]
                if random:
                    t = (philox.uniform_from_index(idx[sp], np.zeros(sp.size, np.int64), seed, 1) * d32).astype(F32)
                else:
                    t = ((ip[sp] / (ip[sp] + F32(1))).astype(F32) * d32).astype(F32)
                t = ((t + starts[c]).astype(F32) * F32(1e6)).astype(F32)
                with np.errstate(invalid='ignore'):
                    ts_all.append(t.astype(np.int64))
                x_all.append((sp % W).astype(np.int16)); y_all.append((sp // W).astype(np.int16))
                p_all.append(np.full(sp.size, 1 - p, np.int8))
        ts = np.concatenate(ts_all)
        order = np.argsort(ts, kind='stable')
        rec = np.empty(ts.size, dtype=EVENT_DTYPE)
        rec['timestamp'] = ts[order]
        rec['x'] = np.concatenate(x_all)[order]
        rec['y'] = np.concatenate(y_all)[order]
        rec['polarity'] = np.concatenate(p_all)[order]
        out.append(rec.view(np.recarray))
    return out
