"""Host-side ceiling of the multi-GPU e2e leg, measured instead of asserted: every rank copies a device
buffer into page-locked host memory at the same time, as the ranks of `bench.py --gpus N` do with their event shards.

    torchrun --nproc-per-node N tools/d2h_ceiling.py [MB per copy = 160] [reps = 20]

Prints one JSON line on rank 0: per-rank and aggregate GB/s for (a) private pinned buffers (cudaHostAlloc),
(b) one shared-memory array mapped and page-locked by every rank (dist.SharedHostRing, what the e2e leg uses),
(c) a single rank alone (the 1-GPU figure), plus the NUMA / CPU affinity each rank ran with."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from v2ce_toolbox_b200 import dist as vdist


def timed_copies(dst, src, reps, barrier):
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return src.numel() * reps / dt / 1e9


def main():
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    aff = vdist.bind_to_gpu_numa(local) if world > 1 else None
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29547')
    vdist.init_process_group('nccl', device=dev, rank=rank, world_size=world)
    try:
        n = mb << 20
        src = torch.randint(0, 255, (n,), dtype=torch.uint8, device=dev)
        barrier = lambda: (dist.barrier(), torch.cuda.synchronize())

        def gather(v):
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return [float(x.item()) for x in out]

        private = gather(timed_copies(torch.empty(n, dtype=torch.uint8, pin_memory=True), src, reps, barrier))
        shared = None
        if world > 1:
            try:
                ring = vdist.SharedHostRing(1, world * n)
                view = ring.buf[rank * n:(rank + 1) * n]
                shared = gather(timed_copies(view, src, reps, barrier))
                ring.close()
            except Exception as e:      # noqa: BLE001
                shared = f'{type(e).__name__}: {e}'
        # one rank at a time: what a single GPU gets from the same host
        alone = []
        for r in range(world):
            v = timed_copies(torch.empty(n, dtype=torch.uint8, pin_memory=True), src, reps, lambda: None) if r == rank else 0.0
            barrier()
            alone.append(v)
        alone = [max(col) for col in zip(*[gather(a) for a in [alone[rank]]])] if False else gather(alone[rank])
        if rank == 0:
            print(json.dumps({'n_gpus': world, 'mb_per_copy': mb, 'reps': reps,
                              'private_pinned_gbs_per_rank': private, 'private_pinned_gbs_total': sum(private),
                              'shared_registered_gbs_per_rank': shared,
                              'shared_registered_gbs_total': sum(shared) if isinstance(shared, list) else None,
                              'one_rank_at_a_time_gbs': alone,
                              'cpu_affinity_rank0': None if aff is None else [aff[0], aff[-1], len(aff)],
                              'host_cpus': os.cpu_count()}), flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
