"""D2H of a large event stream: tensor.cpu() vs sink.to_host.  python tools/sink_bench.py [GB]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from v2ce_toolbox_b200 import sink
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
t = torch.empty(int(gb * (1 << 30)), dtype=torch.uint8, device='cuda').random_(0, 256)
torch.cuda.synchronize()
for name, fn in (('tensor.cpu()', lambda: t.cpu().numpy()), ('sink.to_host', lambda: sink.to_host(t)), ('sink.to_host', lambda: sink.to_host(t)),
                 ('sink 16 workers', lambda: sink.to_host(t, workers=16)), ('sink 256MB chunks', lambda: sink.to_host(t, chunk_bytes=256 << 20))):
    t0 = time.perf_counter(); a = fn(); dt = time.perf_counter() - t0
    print(f'{name:20s} {dt:6.2f} s  {gb / dt:6.2f} GiB/s', flush=True)
    del a
