#!/bin/bash
# Round 2, third GPU call: second-generation one-sweep scatter (ballot ranking), shuffle pack; head conv + pano profiles.
set -u
export PYTHONUNBUFFERED=1
OUT=gpurun_out/r2c
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_ldati.py tests/test_gpu_ldati_pooling.py tests/test_gpu_torch_reference.py tests/test_gpu_pipeline.py tests/test_gpu_event_frames.py -m gpu -q 2>&1 | tail -40 > "$OUT/pytest_gpu.txt"
tail -8 "$OUT/pytest_gpu.txt"
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
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'osw_scatter_kernel|osw_hist_kernel|pack_shfl_kernel' \
    -s 4 -c 4 -o "$OUT/ldati_full_rand" python tools/ldati_bench.py 1 --pairs 24 --dist rand > "$OUT/ncu_ldati_rand.log" 2>&1
# head conv (prep + first depth-merged launch) and the two full-resolution conv2 layers: --set full, one forward
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'head_prep_kernel|conv_halo_kdm_kernel' -c 12 \
    -o "$OUT/kdm_full" python tools/layer_times.py 4 0 > "$OUT/ncu_kdm.log" 2>&1
timeout 120 python tools/layer_times.py 4 5 > "$OUT/layer_times.txt" 2>&1
# where a pano clip spends its host time
timeout 200 python - > "$OUT/pano_profile.txt" 2>&1 <<'PY'
import cProfile, pstats, sys, time
sys.path.insert(0, '.')
import torch
import synth_inputs as synth
import bench
from v2ce_toolbox_b200 import v2ce as drv
dev = torch.device('cuda:0')
n = 200
reader = synth.SynthVideoReader(n, 1080, 1920, seed=0, repeat=4)
reader.cache_range(0, n)
model = bench.new_model(dev)
kw = dict(vidcap=reader, infer_type='pano', seq_len=16, width=346, height=260, batch_size=1, fps=30, seed=1, device=dev,
          write_event_frames=False, events_to_host=False)
drv.stream_clip(model, **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
res = drv.stream_clip(model, **kw)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f'pano 2 tiles, {n} frames: {dt:.3f} s = {(n - 1) / dt:.1f} pairs/s')
pr = cProfile.Profile()
pr.enable()
drv.stream_clip(model, **kw)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(30)
PY
head -3 "$OUT/pano_profile.txt"
ls -la "$OUT"
