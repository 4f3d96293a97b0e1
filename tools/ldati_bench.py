"""LDATI-only microbench (BASELINE configs[2], bounded): python tools/ldati_bench.py [reps] [--variants]
-> bench.ldati_microbench table; --variants repeats it for every combination of the opt-in kernel switches of
csrc/ldati.cu (V2CE_LDATI_REUSE_WARP_TOTALS, V2CE_LDATI_STAGED_SCATTER; read per call)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

args = [a for a in sys.argv[1:] if not a.startswith('--')]
reps = int(args[0]) if args else 5
dev, hbm = torch.device('cuda:0'), bench.peaks()[1]
if '--variants' in sys.argv:
    out = {}
    for reuse in (0, 1):
        for staged in (0, 1):
            os.environ['V2CE_LDATI_REUSE_WARP_TOTALS'] = str(reuse)
            os.environ['V2CE_LDATI_STAGED_SCATTER'] = str(staged)
            r = bench.ldati_microbench(dev, hbm, reps=reps)
            out[f'reuse{reuse}_staged{staged}'] = {k: {'ms': round(v['ms'], 4), 'mevents_per_s': round(v['mevents_per_s'], 1)}
                                                   for k, v in r.items()}
    print(json.dumps(out, indent=1))
else:
    print(json.dumps(bench.ldati_microbench(dev, hbm, reps=reps), indent=1))
