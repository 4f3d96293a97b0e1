"""Only the sharded-vs-single-process equality check of bench.py (no timing): torchrun --nproc-per-node N tools/parity_only.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                              # noqa: E402
from v2ce_toolbox_b200 import dist as vdist               # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    device = torch.device('cuda', local)
    torch.cuda.set_device(device)
    vdist.init_process_group('nccl', device=device)
    out = bench.check_sharded_parity(device, rank, world)
    if rank == 0:
        print(json.dumps({'n_gpus': world, 'sharded_parity': True, 'cases': out}))
    vdist.close_host_group()
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
