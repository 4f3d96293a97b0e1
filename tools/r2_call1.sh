#!/bin/bash
# Round 2, first GPU call: the GPU suite without xfail marks (incl. the torch-CUDA pin of the LDATI `cuda` flavour and the
# image-folder tiny clips), the restructured bench line, and the profiles the kernel work of this round starts from.
#   gpurun --timeout 1200 -- 'bash tools/r2_call1.sh'
set -u
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r2
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > "$OUT/gpu.txt" 2>&1

timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > "$OUT/pytest_gpu.txt"
tail -5 "$OUT/pytest_gpu.txt"

timeout 420 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
tail -c 600 "$OUT/bench.json"; tail -5 "$OUT/bench.err"

# launch list of one dense and one uniform LDATI call (per-kernel times), then --set full of the top kernels
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/ldati_launches_rand.csv" \
    python tools/ldati_bench.py 1 --pairs 24 --dist rand > "$OUT/ldati_rand.log" 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/ldati_launches_randint.csv" \
    python tools/ldati_bench.py 1 --pairs 24 --dist randint10 > "$OUT/ldati_randint.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'sort_scatter_kernel|emit_kernel|pack_kernel|sort_hist_kernel|count_kernel' \
    -s 9 -c 9 -o "$OUT/ldati_full_rand" python tools/ldati_bench.py 1 --pairs 24 --dist rand > "$OUT/ncu_ldati.log" 2>&1
# event frames: launch list + full capture
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'accumulate_kernel|select_hist|select_step|normalize_kernel' \
    -c 12 -o "$OUT/ef_full" python -c "
import sys; sys.path.insert(0, '.')
import torch, bench
print(bench.ef_record(torch.device('cuda:0'), 6531.6, reps=1))" > "$OUT/ncu_ef.log" 2>&1

timeout 180 python tools/cli_e2e.py 321 > "$OUT/cli_e2e.txt" 2>&1
grep -E "frame-pairs/s" "$OUT/cli_e2e.txt"
timeout 120 python tools/layer_times.py 4 5 > "$OUT/layer_times.txt" 2>&1
head -3 "$OUT/layer_times.txt"
ls -la "$OUT"
