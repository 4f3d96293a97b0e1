#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out gets its first
# hardware run, and the profiles the next optimisation steps need are captured in one go.
#
#   gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
#
# Outputs land in gpurun_out/r2/ (merged back by gpurun).  Every step has its own timeout; none depends on another.
set -u
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r2
mkdir -p "$OUT"

# 1. the whole GPU suite, verbose for the xfail-marked newcomers (XPASS = they work): device-side resize
#    (tests/test_gpu_preprocess.py), LDATI pooling (tests/test_gpu_ldati_pooling.py), the TF32 probe
timeout 300 python -m pytest tests -m gpu -q -rxX 2>&1 | tail -40 > "$OUT/pytest_gpu.txt"
tail -5 "$OUT/pytest_gpu.txt"

# 2. LDATI microbench, all kernel-variant combinations (CUDA events)
timeout 120 python tools/ldati_bench.py 5 --variants > "$OUT/ldati_variants.json" 2>&1

# 3. ncu --set full of the two LDATI kernels that dominate: one launch each from the dense 24-pair call
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'sort_scatter_kernel|emit_kernel' -s 6 -c 3 \
    -o "$OUT/ldati_full" python tools/ldati_bench.py 1 > "$OUT/ncu_ldati.log" 2>&1

# 4. ncu --set full of the head conv (prep + depth-merged kernel): the largest gap left in the forward
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'head_prep_kernel|conv_halo_kdm_kernel' -s 0 -c 2 \
    -o "$OUT/head_full" python tools/layer_times.py 4 1 > "$OUT/ncu_head.log" 2>&1

# 5. the drop-in CLI end to end with the new host side (prefetch, background batches, threaded encode, npz sink)
timeout 180 python tools/cli_e2e.py 321 > "$OUT/cli_e2e.txt" 2>&1
timeout 120 python tools/sink_bench.py 2 > "$OUT/sink_bench.txt" 2>&1
grep -E "frame-pairs/s" "$OUT/cli_e2e.txt"

# 6. per-layer times and the bench line
timeout 120 python tools/layer_times.py 4 5 > "$OUT/layer_times.txt" 2>&1
timeout 240 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
tail -c 300 "$OUT/bench.json"
