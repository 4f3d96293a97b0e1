#!/bin/bash
# Round 2, sixth GPU call: kdm issue loop with compile-time K steps; CLI stage times; pano retime.
set -u
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r2f
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_pipeline.py tests/test_gpu_torch_reference.py tests/test_gpu_ldati.py -m gpu -q -x 2>&1 | tail -30 > "$OUT/pytest_gpu.txt"
tail -6 "$OUT/pytest_gpu.txt"
timeout 120 python tools/layer_times.py 4 5 > "$OUT/layer_times.txt" 2>&1; cat "$OUT/layer_times.txt"
timeout 300 python bench.py --headline-only > "$OUT/bench_headline.json" 2> "$OUT/bench.err"; python -c "
import json
d=json.loads(open('$OUT/bench_headline.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'fwd', d['roofline']['forward_ms'], 'frac', d['roofline']['frac'], d['roofline']['frac_burst'], d['clocks'])"
timeout 180 python tools/cli_e2e.py 321 > "$OUT/cli_e2e.txt" 2>&1; grep -E "frame-pairs/s" "$OUT/cli_e2e.txt"
