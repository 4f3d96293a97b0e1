#!/bin/bash
# ncu: full capture of the kdm launches of one forward after the edge-slice peeling; launch list of the bench command.
set -u
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r2j
mkdir -p "$OUT"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv_halo_kdm_kernel' -c 11 \
    -o "$OUT/kdm_full" python tools/layer_times.py 4 0 > "$OUT/ncu_kdm.log" 2>&1
tail -2 "$OUT/ncu_kdm.log"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches_bench.csv" \
    python bench.py --headline-only --steps 2 --warmup 1 > "$OUT/bench_under_ncu.log" 2>&1
tail -c 300 "$OUT/bench_under_ncu.log"; wc -l "$OUT/launches_bench.csv"
