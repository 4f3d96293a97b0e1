"""Driver-level restatement: frames -> windows -> voxels -> event frames + event stream.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows /root/reference/v2ce.py:

  preprocess        <- v2ce.py:45-64   (image_pre_processing)
  window_starts     <- v2ce.py:149-154 (sequence_num / mode / pulled-back last window)
  center_crop       <- v2ce.py:78
  pano_tiles        <- v2ce.py:103-111,121-126
  video_to_voxels   <- v2ce.py:132-239 (incl. merge_voxels)
  event_stream      <- v2ce.py:351-367 (stage-2 chunking and per-frame time offset)
"""
import numpy as np
import torch

from . import ldati_oracle


def preprocess(images, height=260):
    """images (N,H,W) uint8 -> image units (N-1,2,H',W') float32."""
    import cv2
    im = images.astype(np.float32) / 255
    im = np.stack([cv2.resize(a, (int(a.shape[1] / a.shape[0] * height), height)) for a in im], axis=0)
    units = np.stack([im[:-1], im[1:]], axis=1)
    return ((units - np.float32(0.153)) / np.float32(0.165)).astype(np.float32)


def window_starts(frame_count, seq_len=16):
    n_seq = int(np.ceil((frame_count - 1) / seq_len))
    mode = (frame_count - 1) % seq_len
    starts = np.arange(n_seq) * seq_len
    if mode != 0:
        starts[-1] -= (seq_len - mode)
    return starts, mode


def center_crop(units, width=346):
    c = units.shape[-1] // 2
    return units[..., c - width // 2:c + width // 2]


def pano_tiles(total_width, width=346):
    """[(src_start, src_end, keep_last_n_columns)] per tile."""
    n = int(np.ceil(total_width / width))
    exact = total_width % 346 == 0          # sic: the reference tests against the literal 346 (v2ce.py:104)
    rem = total_width % width
    tiles = []
    for i in range(n):
        if i == n - 1 and not exact:
            tiles.append((total_width - width, total_width, rem))
        else:
            tiles.append((i * width, (i + 1) * width, width))
    return tiles


def video_to_voxels(model_forward, frames, infer_type='center', seq_len=16, width=346, height=260,
                    batch_size=1, read_frame=None):
    """model_forward: callable (B,L,2,H,W) tensor -> (B,L,20,H,W) tensor, called in the
    reference's order (per batch; per tile inside a batch for pano)."""
    frame_count = len(frames)
    starts, mode = window_starts(frame_count, seq_len)
    pending, outs = [], []
    for si, st in enumerate(starts):
        idx = range(st, st + seq_len + 1)
        imgs = np.stack([frames[max(i, 0)] for i in idx], axis=0)   # index -1 (16-frame clips) reads frame 0, as cv2 does
        pending.append(preprocess(imgs, height)[None])
        if len(pending) == batch_size or si == len(starts) - 1:
            batch = torch.from_numpy(np.concatenate(pending, axis=0))
            pending = []
            if infer_type == 'center':
                out_w = width
                pred = model_forward(center_crop(batch, width))
            else:
                out_w = batch.shape[-1]
                parts = []
                for (a, b, keep) in pano_tiles(batch.shape[-1], width):
                    o = model_forward(batch[..., a:b])
                    parts.append(o[..., -keep:] if keep != width else o)
                pred = torch.cat(parts, dim=-1)
            outs.append(pred.numpy())
    return merge_voxels(outs, height, out_w, mode)


def merge_voxels(outs, height, width, mode):
    """outs: list of (b,L,20,H,W) -> (N,2,10,H,W); the pulled-back last window only
    contributes its last `mode` pairs."""
    chunks = []
    for o in outs[:-1]:
        chunks.append(o.reshape(-1, 2, 10, height, width))
    last = outs[-1]
    if last.shape[0] > 1:
        chunks.append(last[:-1].reshape(-1, 2, 10, height, width))
    tail = last[-1][-mode:] if mode != 0 else last[-1]
    chunks.append(tail.reshape(-1, 2, 10, height, width))
    return np.concatenate(chunks, axis=0)


def frame_offset_us(i, fps):
    return int(i * 1 / fps * 1e6)


def event_stream(voxels, fps=30, stage2_batch_size=24, seed=0, flavor='cuda'):
    """Whole-clip event stream with per-frame offsets (v2ce.py:359-367)."""
    recs = []
    for i in range(0, voxels.shape[0], stage2_batch_size):
        recs.extend(ldati_oracle.sample_voxel_statistical_oracle(
            voxels[i:i + stage2_batch_size], fps=fps, seed=seed, frame_base=i, flavor=flavor))
    for i, r in enumerate(recs):
        r['timestamp'] += frame_offset_us(i, fps)
    return np.concatenate(recs)
