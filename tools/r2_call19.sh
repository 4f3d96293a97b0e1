#!/bin/bash
# Round 2 closing run on one GPU: full GPU suite, smoke, sanitizers, the full bench line, the reference arm.
set -u
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r2k
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > "$OUT/pytest_gpu.txt"; tail -4 "$OUT/pytest_gpu.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
bash tools/sanitize.sh "$OUT/sanitize" 2>&1 | tail -6
timeout 1500 python bench.py > "$OUT/bench_final.json" 2> "$OUT/bench_final.err"; echo "bench rc $?"; tail -3 "$OUT/bench_final.err"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "ref rc $?"; tail -c 600 "$OUT/bench_reference.json"
python - <<PY
import json
for l in open('$OUT/bench_final.json'):
    l = l.strip()
    if not l.startswith('{'): continue
    d = json.loads(l)
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e'])
    print({k: d['roofline'].get(k) for k in ('frac', 'frac_burst', 'forward_ms', 'forward_ms_in_step')}, d['clocks'])
    print('ef', d.get('ef')); print('lib', json.dumps(d.get('library_baseline'))[:700]); print('cpu', d.get('cpu_baseline'))
    print('ldati', json.dumps(d.get('ldati'))[:900])
    print('clips', json.dumps(d.get('clips'))[:900])
PY
