#!/bin/bash
# Round 2, fourth GPU call: LDATI with the private-counter scatter + linear pack; kdm ncu summary; smoke.
set -u
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r2d
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_ldati.py tests/test_gpu_ldati_pooling.py tests/test_gpu_torch_reference.py tests/test_gpu_pipeline.py tests/test_gpu_event_frames.py tests/test_gpu_unet.py -m gpu -q 2>&1 | tail -40 > "$OUT/pytest_gpu.txt"
tail -8 "$OUT/pytest_gpu.txt"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.txt" 2>&1; tail -2 "$OUT/smoke.txt"
timeout 300 python tools/ldati_bench.py 5 --table > "$OUT/ldati_table.json" 2> "$OUT/ldati_table.err"
python -c "
import json
d=json.load(open('$OUT/ldati_table.json'))
for k,v in d.items(): print(k, round(v['ms'],3), 'ms', round(v['frac_of_hbm_peak'],3))
"
for d in rand randint10 sparse; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/ldati_launches_$d.csv" \
      python tools/ldati_bench.py 1 --pairs 24 --dist $d > "$OUT/ldati_$d.log" 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'pack_linear_kernel|osw_hist_kernel' \
    -s 2 -c 2 -o "$OUT/ldati_full_rand" python tools/ldati_bench.py 1 --pairs 24 --dist rand > "$OUT/ncu_ldati_rand.log" 2>&1
ls -la "$OUT"
