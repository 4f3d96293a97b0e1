#!/bin/bash
# Round 2: first hardware run of the baseline samplers and the ts_diff metric; full GPU suite; sanitizer pass.
set -u
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r2g
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_baseline.py tests/test_gpu_metric.py -m gpu -q 2>&1 | tail -40 > "$OUT/pytest_new.txt"; tail -12 "$OUT/pytest_new.txt"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > "$OUT/pytest_gpu.txt"; tail -4 "$OUT/pytest_gpu.txt"
bash tools/sanitize.sh "$OUT/sanitize" 2>&1 | tail -6
