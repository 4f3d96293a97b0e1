"""Experiment: one batch-4 forward vs two concurrent batch-2 forwards on two streams (tail filling)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synth_inputs as synth
from v2ce_toolbox_b200.scripts.v2ce_3d import V2ce3d

H, W, L = 260, 346, 16
def mk():
    m = V2ce3d(); m.load_state_dict(synth.make_state_dict(0, 'reference')); return m.eval().to('cuda:0')
m, m1, m2 = mk(), mk(), mk()
x = torch.randn(4, L, 2, H, W, device='cuda:0')
xa, xb = x[:2].contiguous(), x[2:].contiguous()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def single():
    m(x)
def split():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): m1(xa)
    with torch.cuda.stream(s2): m2(xb)
    cur.wait_stream(s1); cur.wait_stream(s2)
for name, fn in (('single', single), ('split', split), ('single', single), ('split', split)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 10, 'ms')
