#!/bin/bash
# Round 2, 2-GPU call: NCCL tests, the bench line at N = 2 (sharded_parity, shared host ring, clips), D2H ceiling.
set -u
export PYTHONUNBUFFERED=1
N=${1:-2}
OUT=gpurun_out/r2_${N}gpu
mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -15 > "$OUT/pytest_dist.txt"; tail -4 "$OUT/pytest_dist.txt"
fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    tools/d2h_ceiling.py 160 20 > "$OUT/d2h_ceiling.json" 2> "$OUT/d2h_ceiling.err"; tail -c 900 "$OUT/d2h_ceiling.json"; tail -3 "$OUT/d2h_ceiling.err"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"
tail -5 "$OUT/bench.err"
python - <<PY
import json
for l in open('$OUT/bench.json'):
    l = l.strip()
    if not l.startswith('{'): continue
    d = json.loads(l)
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'sharded_parity')}, d['e2e'])
    print(json.dumps(d.get('clips'))[:1500])
    print(d['config']['multi_gpu'])
    print(d['roofline']['forward_ms'], d['roofline']['forward_ms_in_step'])
PY
