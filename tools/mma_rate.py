import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from v2ce_toolbox_b200 import _lib
lib = _lib.load()
torch.cuda.init(); torch.zeros(1, device='cuda')
for ctas in (1, 148):
    for bn in (32, 64, 96, 128, 192, 256):
        for naccs in (1, 2):
            if naccs * bn > 512: continue
            c = ctypes.c_double()
            _lib.check(lib.v2ce_debug_mma_rate(bn, 4000, naccs, ctas, ctypes.byref(c)))
            print(f'ctas={ctas:3d} N={bn:3d} accs={naccs}: {c.value:7.1f} cycles/MMA  (math floor {bn//2})')
