"""BASELINE configs[3] / configs[4]: whole clips through the sharded CLI path (dist.stream_clip_sharded), one process per GPU.

    torchrun --nproc-per-node N tools/clip_bench.py long [n_frames=9000]       # center 346x260, windows sharded, NCCL event merge
    torchrun --nproc-per-node N tools/clip_bench.py pano [n_frames=600]        # 1920x1080 -> 462x260 (host cv2 resize), 2 width tiles

Prints one JSON line on rank 0: wall-clock of the whole job (frame synthesis excluded), per-stage seconds, frame-pairs/s and
events/s.  Frames are synthetic (oracle/synth.make_video), weights random-init (seed 0).
"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import synth_inputs as synth
from synth_inputs import FakeVideoReader
from v2ce_toolbox_b200 import dist as vdist
from v2ce_toolbox_b200.scripts.v2ce_3d import V2ce3d


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else 'long'
    n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else (9000 if kind == 'long' else 600)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        from v2ce_toolbox_b200.dist import bind_to_gpu_numa
        bind_to_gpu_numa(local)                  # pinned staging and copy threads on the GPU's own NUMA node
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29533')
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        if kind == 'long':
            frames = synth.make_video(n_frames, 260, 346, seed=0)
            kw = dict(infer_type='center', width=346, height=260)
        else:
            # a 1080p source; a small texture tiled up keeps the synthesis cheap
            small = synth.make_video(n_frames, 270, 480, seed=0)
            frames = np.repeat(np.repeat(small, 4, axis=1), 4, axis=2)          # (n, 1080, 1920)
            kw = dict(infer_type='pano', width=346, height=260)
        model = V2ce3d()
        model.load_state_dict(synth.make_state_dict(0, 'reference'))
        model.eval().to(dev)
        reader = FakeVideoReader(frames)
        # warm-up on a short prefix (library load, first-touch allocations, NCCL communicators)
        vdist.stream_clip_sharded(model, FakeVideoReader(frames[:17 * world + 1]), 17 * world + 1, world, rank, seq_len=16,
                                  batch_size=4, fps=30, seed=1, device=dev, **kw)
        model2 = V2ce3d()
        model2.load_state_dict(synth.make_state_dict(0, 'reference'))
        model2.eval().to(dev)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev_dev, n_events = vdist.stream_clip_sharded(model2, reader, n_frames, world, rank, seq_len=16, batch_size=4, fps=30,
                                                     seed=1, device=dev, to_host=False, **kw)
        torch.cuda.synchronize()
        dist.barrier()
        dt_dev = time.perf_counter() - t0             # every rank computed, shards merged on rank 0's device
        if rank == 0:
            from v2ce_toolbox_b200.ldati import EVENT_DTYPE
            from v2ce_toolbox_b200.sink import to_host
            ev = to_host(ev_dev).view(EVENT_DTYPE)       # the event-stream sink: pipelined D2H into one host array
            dt = time.perf_counter() - t0
            ts = ev['timestamp']
            ok = bool((np.diff(ts[::max(1, len(ts) // 2000000)]) >= -40000).all())      # bins restart every 1/fps/9 inside a frame
            print(json.dumps({'workload': kind, 'n_frames': n_frames, 'n_gpus': world, 'pairs': n_frames - 1,
                              'events': int(n_events), 'wall_s': dt, 'device_s': dt_dev, 'pairs_per_s': (n_frames - 1) / dt,
                              'pairs_per_s_device': (n_frames - 1) / dt_dev,
                              'mevents_per_s': n_events / dt / 1e6, 'stream_bytes': int(n_events) * 13,
                              'frame_shape': list(frames.shape[1:]), 'monotone_frames': ok}), flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
