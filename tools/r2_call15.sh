#!/bin/bash
# A/B on ONE box: kdm issue loop with the edge slices peeled (new) against the build before it (build/ab/lib_before.so).
set -u
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r2i
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_pipeline.py tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -5 > "$OUT/pytest_gpu.txt"; tail -3 "$OUT/pytest_gpu.txt"
run() { name=$1; timeout 120 python tools/layer_times.py 4 8 > "$OUT/layer_times_$name.txt" 2>&1; echo "== $name: $(head -1 $OUT/layer_times_$name.txt) | $(tail -1 $OUT/layer_times_$name.txt)"; }
run new
timeout 300 python bench.py --headline-only --steps 20 --warmup 3 > "$OUT/bench_headline.json" 2> "$OUT/bench_headline.err"; tail -c 600 "$OUT/bench_headline.json"
if [ -f build/ab/lib_before.so ]; then
  cp v2ce_toolbox_b200/libv2ce_b200.so /tmp/lib_new.so
  cp build/ab/lib_before.so v2ce_toolbox_b200/libv2ce_b200.so
  run before
  cp /tmp/lib_new.so v2ce_toolbox_b200/libv2ce_b200.so
  run new_again
fi
