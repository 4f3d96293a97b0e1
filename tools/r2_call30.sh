#!/bin/bash
# Spectral-norm kernels rewritten for memory-level parallelism: tests, time per replayed step, forward.
set -u
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r2q
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_pipeline.py tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -4
timeout 200 python - <<'PY'
import torch, bench
m = bench.new_model(torch.device('cuda', 0))
m.sn_advance(5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m.sn_advance(200); e1.record(); torch.cuda.synchronize()
print('sn step: %.1f us' % (e0.elapsed_time(e1) / 200 * 1e3))
PY
timeout 120 python tools/layer_times.py 4 8 > "$OUT/layer_times.txt" 2>&1; head -1 "$OUT/layer_times.txt"; tail -1 "$OUT/layer_times.txt"
timeout 300 python bench.py --headline-only --steps 30 --warmup 3 > "$OUT/bench_headline.json" 2> "$OUT/bench.err"
python - <<PY
import json
for l in open("$OUT/bench_headline.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print(round(d["value"]), round(d["ms_per_step"], 3), round(r["forward_ms"], 3), round(r["forward_ms_in_step"], 3), round(d["e2e"]["value"]))
PY
