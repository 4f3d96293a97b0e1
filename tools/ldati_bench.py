"""LDATI-only microbench (BASELINE configs[2], bounded): python tools/ldati_bench.py  -> bench.ldati_microbench table."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

print(json.dumps(bench.ldati_microbench(torch.device('cuda:0'), bench.peaks()[1], reps=int(sys.argv[1]) if len(sys.argv) > 1 else 5), indent=1))
