"""LDATI-only runs for profiling: python tools/ldati_bench.py [reps] [--pairs F] [--dist rand|randint10|sparse] [--table]

  --table      the whole bench.ldati_microbench table (24 / 96 / 1000 pairs), as bench.py reports it
  otherwise    `reps` count -> emit calls on F pairs of one distribution (what `ncu -k regex:...` is pointed at), then
               per-kernel device times of one call from CUDA events are NOT taken here: use the ncu launch list."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench


def arg(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


pos = [a for i, a in enumerate(sys.argv[1:], 1) if not a.startswith('--') and not sys.argv[i - 1].startswith('--')]
reps = int(pos[0]) if pos else 5
dev, hbm = torch.device('cuda:0'), bench.peaks()['hbm']
if '--table' in sys.argv:
    print(json.dumps(bench.ldati_microbench(dev, hbm, reps=reps), indent=1))
else:
    from v2ce_toolbox_b200 import ldati
    F, name = int(arg('--pairs', 24)), arg('--dist', 'rand')
    g = torch.Generator(device=dev).manual_seed(42)
    if name == 'rand':
        vox = torch.rand((F, 2, 10, bench.H, bench.W), generator=g, device=dev)
    elif name == 'sparse':
        vox = torch.rand((F, 2, 10, bench.H, bench.W), generator=g, device=dev) * 0.015
    else:
        vox = torch.randint(0, 10, (F, 2, 10, bench.H, bench.W), generator=g, device=dev).float()
    eng = ldati.LdatiEngine(dev)
    params = ldati.make_params(F, bench.H, bench.W, fps=30, seed=42, frame_base=0, device=dev)
    total = [0]

    def run():
        _, seg, _ = eng.run(vox, params)
        total[0] = int(seg.sum())
    ms = bench._time_ms(run, reps, warm=1)
    print(json.dumps({'dist': name, 'pairs': F, 'events': total[0], 'ms': ms, 'mevents_per_s': total[0] / ms / 1e3}))
