#!/bin/bash
# Closing run: GPU suite, smoke, full bench line, CLI wall clock with the parallel .npz writer.
set -u
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r2o
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > "$OUT/pytest_gpu.txt"; tail -3 "$OUT/pytest_gpu.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py > "$OUT/bench_final.json" 2> "$OUT/bench_final.err"; echo "bench rc $?"; tail -3 "$OUT/bench_final.err"
timeout 300 python tools/cli_e2e.py 321 > "$OUT/cli_e2e.txt" 2>&1; head -2 "$OUT/cli_e2e.txt"
python - <<PY
import json
for l in open('$OUT/bench_final.json'):
    l = l.strip()
    if not l.startswith('{'): continue
    d = json.loads(l)
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e'])
    print({k: d['roofline'].get(k) for k in ('frac', 'frac_burst', 'forward_ms', 'forward_ms_in_step', 'forward_gap_ms_in_step')}, d['clocks'])
    for k, v in d['clips'].items():
        if isinstance(v, dict): print(k, {a: (round(b['s'], 3), round(b['pairs_per_s'])) for a, b in v.items() if isinstance(b, dict)})
PY
