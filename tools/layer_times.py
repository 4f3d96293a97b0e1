"""Per-launch device times of one V2ce3d forward (CUDA events on the launch stream, warm caches).

    python tools/layer_times.py [batch] [reps]  ->  table: layer, ms, GFLOP, TFLOP/s, share
Used to decide which kernel to work on; the numbers behind profiles/*layer_times*.txt.
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import synth_inputs as synth
from v2ce_toolbox_b200.scripts.v2ce_3d import V2ce3d

H, W, L = 260, 346, 16


def layer_gflop(B):
    """2*MACs per layer at 346x260, keyed by layer name."""
    hs, ws = [H], [W]
    for _ in range(4):
        hs.append((hs[-1] - 1) // 2 + 1)
        ws.append((ws[-1] - 1) // 2 + 1)
    ch = [32, 64, 128, 256, 512]
    out = {'UNet.head.conv3d': 2 * B * L * H * W * 32 * 2 * 27}
    for i in range(4):
        m = B * L * hs[i + 1] * ws[i + 1]
        out[f'UNet.encoders.{i}.conv1'] = 2 * m * ch[i + 1] * ch[i] * 27
        out[f'UNet.encoders.{i}.conv2'] = 2 * m * ch[i + 1] * ch[i + 1] * 27
        out[f'UNet.encoders.{i}.downsample.0'] = 2 * m * ch[i + 1] * ch[i]
    m = B * L * hs[4] * ws[4]
    for i in range(2):
        out[f'UNet.resblocks.{i}.conv1'] = out[f'UNet.resblocks.{i}.conv2'] = 2 * m * 512 * 512 * 27
        out[f'UNet.resblocks.{i}.downsample.0'] = 2 * m * 512 * 512
    for i in range(4):
        lvl = 3 - i
        m = B * L * hs[lvl] * ws[lvl]
        cin = ch[lvl + 1] + ch[lvl]
        out[f'UNet.decoders.{i}.conv1'] = 2 * m * ch[lvl] * cin * 27
        out[f'UNet.decoders.{i}.conv2'] = 2 * m * ch[lvl] * ch[lvl] * 27
        out[f'UNet.decoders.{i}.downsample.0'] = 2 * m * ch[lvl] * cin
    out['UNet.decoders.3.conv2'] += 2 * B * L * H * W * 20 * 32       # fused prediction layer
    return {k: v / 1e9 for k, v in out.items()}


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    model = V2ce3d()
    model.load_state_dict(synth.make_state_dict(0, 'reference'))
    model.eval().to('cuda:0')
    model.set_option('layer_timing', 1)
    x = torch.randn(B, L, 2, H, W, device='cuda:0')
    acc = None
    for r in range(reps + 2):
        model(x)
        t = model.layer_times()
        if r >= 2:
            acc = [(n, a + ms) for (n, ms), (_, a) in zip(t, acc)] if acc else t
    gf = layer_gflop(B)
    seen = {n for n, _ in acc}
    for i in range(4):                              # shortcut fused into conv1: count its FLOPs there
        sc = f'UNet.decoders.{i}.downsample.0'
        if sc not in seen:
            gf[f'UNet.decoders.{i}.conv1'] += gf.pop(sc)
    tot = sum(ms for _, ms in acc) / reps
    print(f'batch {B}, {reps} reps, forward {tot:.3f} ms = {sum(gf.values()) / tot:.1f} TFLOP/s')
    print(f'{"layer":34s} {"ms":>8s} {"GFLOP":>9s} {"TFLOP/s":>9s} {"share":>6s}')
    for n, ms in acc:
        ms /= reps
        print(f'{n:34s} {ms:8.3f} {gf[n]:9.1f} {gf[n] / ms:9.1f} {100 * ms / tot:5.1f}%')
    # the forward as the product runs it (no per-launch events; shortcut launches may overlap conv tails)
    model.set_option('layer_timing', 0)
    model(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        model(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f'untimed forward: {ms:.3f} ms = {sum(gf.values()) / ms:.1f} TFLOP/s')


if __name__ == '__main__':
    main()
