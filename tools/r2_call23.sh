#!/bin/bash
# After the N = 128 tiles for the 256-channel layers: GPU suite, smoke, full bench line, bench launch list.
set -u
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r2m
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > "$OUT/pytest_gpu.txt"; tail -3 "$OUT/pytest_gpu.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py > "$OUT/bench_final.json" 2> "$OUT/bench_final.err"; echo "bench rc $?"; tail -3 "$OUT/bench_final.err"
timeout 120 python tools/layer_times.py 4 8 > "$OUT/layer_times.txt" 2>&1; head -1 "$OUT/layer_times.txt"; tail -1 "$OUT/layer_times.txt"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches_bench.csv" \
    python bench.py --headline-only --steps 2 --warmup 1 > "$OUT/bench_under_ncu.log" 2>&1
wc -l "$OUT/launches_bench.csv"
python - <<PY
import json
for l in open('$OUT/bench_final.json'):
    l = l.strip()
    if not l.startswith('{'): continue
    d = json.loads(l)
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e'])
    print({k: d['roofline'].get(k) for k in ('frac', 'frac_burst', 'forward_ms', 'forward_ms_in_step')}, d['clocks'])
    print('lib', json.dumps(d.get('library_baseline'))[:400])
PY
