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

# the .npz leg of the sink (host only): np.savez vs sink.save_npz on the same event array
import tempfile
import numpy as np
from v2ce_toolbox_b200.ldati import EVENT_DTYPE
n_ev = int(min(gb, 2.0) * (1 << 30)) // 13
ev = np.zeros(n_ev, EVENT_DTYPE)
ev['timestamp'] = np.arange(n_ev)
with tempfile.TemporaryDirectory() as d:
    for name, fn in (('np.savez', lambda p: np.savez(p, event_stream=ev)),
                     ('sink.save_npz', lambda p: sink.save_npz(p, event_stream=ev)),
                     ('sink.save_npz 16w', lambda p: sink.save_npz(p, workers=16, event_stream=ev))):
        p = os.path.join(d, name.replace(' ', '_') + '.npz')
        t0 = time.perf_counter(); fn(p); dt = time.perf_counter() - t0
        print(f'{name:20s} {dt:6.2f} s  {ev.nbytes / dt / 1e9:6.2f} GB/s  ({ev.nbytes / 1e9:.2f} GB of events)', flush=True)
        os.remove(p)
