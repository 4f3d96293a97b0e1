"""Wall-clock of the drop-in CLI on BASELINE configs[1]: 321 synthetic 346x260 PNG frames, -b 4, event-frame video on.

    python tools/cli_e2e.py [n_frames]   ->  seconds per stage and frame-pairs/s, including PNG decode (cv2), the
    device pipeline, D2H, the mp4v encode (cv2) and np.savez of the event stream -- everything `python v2ce.py -f ...` does.
"""
import os
import sys
import tempfile
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
import torch
import synth_inputs as synth
from v2ce_toolbox_b200 import v2ce as drv


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 321
    tmp = tempfile.mkdtemp(prefix='v2ce_cli_')
    folder = os.path.join(tmp, 'clip')
    os.makedirs(folder)
    frames = synth.make_video(n, 260, 346, seed=0)
    for i, f in enumerate(frames):
        cv2.imwrite(os.path.join(folder, f'{i:05d}.png'), f)
    ckpt = os.path.join(tmp, 'w.pt')
    torch.save(synth.make_state_dict(0, 'reference'), ckpt)
    out = os.path.join(tmp, 'out')
    argv = ['-f', folder, '-o', out, '-m', ckpt, '-b', '4', '--seed', '1', '-l', 'warning']
    drv.main(argv)                                   # warm-up: library load, first-touch allocations
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = drv.main(argv)
    dt = time.perf_counter() - t0
    pairs = res.n_pairs
    print(f'{n} frames -> {pairs} pairs, {len(res.event_stream)} events, {dt:.3f} s wall = {pairs / dt:.1f} frame-pairs/s '
          f'(PNG decode + device pipeline + D2H + mp4v encode + npz); stages: '
          + ', '.join(f'{k} {v:.3f}' for k, v in res.timings.items()))
    # the same without the host-side codecs: stream_clip on in-memory frames
    from synth_inputs import FakeVideoReader
    model = drv.get_trained_mode(ckpt)
    drv.stream_clip(model, vidcap=FakeVideoReader(frames), batch_size=4, seed=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = drv.stream_clip(model, vidcap=FakeVideoReader(frames), batch_size=4, seed=1)
    dt = time.perf_counter() - t0
    print(f'stream_clip on in-memory frames: {dt:.3f} s wall = {res.n_pairs / dt:.1f} frame-pairs/s')
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    drv.stream_clip(model, vidcap=FakeVideoReader(frames), batch_size=4, seed=1)
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(18)


if __name__ == '__main__':
    main()
