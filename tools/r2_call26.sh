#!/bin/bash
# High-priority network stream inside BatchRunner: GPU suite, then A/B of the bench headline on one box.
set -u
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r2n
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > "$OUT/pytest_gpu.txt"; tail -3 "$OUT/pytest_gpu.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for v in 1 0 1 0 1; do
  V2CE_NET_STREAM=$v timeout 300 python bench.py --headline-only --steps 30 --warmup 3 > "$OUT/bench_net$v.json" 2> "$OUT/bench.err"
  python - <<PY
import json
for l in open("$OUT/bench_net$v.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("net_stream=$v", round(d["value"]), round(d["ms_per_step"], 3), round(r["forward_ms"], 3), round(r["forward_ms_in_step"], 3), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3))
PY
done
