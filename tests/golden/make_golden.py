"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (/root/reference) on CPU in the build container.

    python tests/golden/make_golden.py

The GPU box has no /root/reference, so the parity tests there read these files.
Inputs come from the synth_inputs.py generators (seeded; re-exported as oracle.synth); small inputs are stored
next to the outputs, full-size cases store only counts and SHA-256 digests of
the tie-canonicalised event bytes.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness as rh                     # noqa: E402
from oracle import synth, ldati_oracle as lo            # noqa: E402


def digest(rec):
    c = lo.canonicalize(rec)
    return hashlib.sha256(np.ascontiguousarray(c).tobytes()).hexdigest()


def ldati_cases():
    out = {}
    meta = {}
    y = np.zeros((1, 2, 10, 1, 1), np.float32)
    # the only known-answer vector in the reference repo:
    # train/scripts/stage2/vis_stage2.ipynb cells 1-2 (positive polarity, fps 30)
    y[0, 0, :, 0, 0] = [0, 0, .9179, .0821, .9962, .0038, .5287, 2.8454, 1.6884, .9375]
    small = {
        'kat': (y, 30),
        'rand': (synth.make_voxels('rand', 2, 40, 52, seed=42), 30),
        'randint': (synth.make_voxels('randint', 1, 24, 30, seed=43), 30),
        'sparse': (synth.make_voxels('sparse', 2, 40, 52, seed=44), 30),
        'mixed24': (synth.make_voxels('mixed', 2, 33, 47, seed=45), 24),
        'mixed120': (synth.make_voxels('mixed', 1, 33, 47, seed=46), 120),
    }
    for name, (v, fps) in small.items():
        ref = rh.run_reference_ldati(v, fps=fps, seed=42, frame_base=5)
        out[f'{name}_voxel'] = v
        for i, r in enumerate(ref):
            out[f'{name}_events_{i}'] = np.ascontiguousarray(lo.canonicalize(r)).view(np.uint8)
        meta[name] = dict(fps=fps, seed=42, frame_base=5, frames=len(ref), counts=[int(len(r)) for r in ref])
    # full-size digests (inputs regenerated from the seed)
    for kind, seed in (('rand', 42), ('sparse', 7), ('randint', 9)):
        v = synth.make_voxels(kind, 1, 260, 346, seed=seed)
        ref = rh.run_reference_ldati(v, fps=30, seed=42, frame_base=0)
        meta[f'full_{kind}'] = dict(fps=30, seed=42, frame_base=0, voxel_seed=seed, kind=kind,
                                    counts=[int(len(r)) for r in ref], sha256=[digest(r) for r in ref])
    np.savez_compressed(os.path.join(HERE, 'ldati_golden.npz'), **out)
    return meta


OPTION_CASES = [(inp, strat, bid) for inp in ('mixed24', 'randint') for strat in ('slope', 'random', 'none')
                for bid in (False, True) if not (strat == 'slope' and not bid)]


def ldati_option_cases():
    """The options of sample_voxel_statistical the CLI does not use: additional_events_strategy
    'random' (LDATI.py:173-174) and 'none' (LDATI.py:206-207,241-244), and bidirectional=True (LDATI.py:107-122),
    on the 'mixed24' and 'randint' voxels already stored in ldati_golden.npz."""
    vox = np.load(os.path.join(HERE, 'ldati_golden.npz'))
    with open(os.path.join(HERE, 'golden_meta.json')) as f:
        base = json.load(f)['ldati']
    out, meta = {}, {}
    for inp, strat, bid in OPTION_CASES:
        m = base[inp]
        ref = rh.run_reference_ldati(vox[f'{inp}_voxel'], fps=m['fps'], seed=m['seed'], frame_base=m['frame_base'],
                                     additional_events_strategy=strat, bidirectional=bid)
        name = f"{inp}-{strat}-{'bi' if bid else 'uni'}"
        for i, r in enumerate(ref):
            out[f'{name}_events_{i}'] = np.ascontiguousarray(lo.canonicalize(r)).view(np.uint8)
        meta[name] = dict(input=inp, fps=m['fps'], seed=m['seed'], frame_base=m['frame_base'], frames=len(ref),
                          additional_events_strategy=strat, bidirectional=bid, counts=[int(len(r)) for r in ref])
    np.savez_compressed(os.path.join(HERE, 'ldati_options_golden.npz'), **out)
    return meta


POOLING_CASES = [('weighted', 3, False), ('avg', 3, False), ('avg', 5, True)]


def ldati_pooling_cases():
    """pooling_type 'weighted' / 'avg' (LDATI.py:176-183) on the 'mixed24' voxels of ldati_golden.npz.  Only events of
    multi-event pixel-bins depend on the pooling, so the fixture stores the canonicalised events of frame 0 only."""
    vox = np.load(os.path.join(HERE, 'ldati_golden.npz'))
    with open(os.path.join(HERE, 'golden_meta.json')) as f:
        m = json.load(f)['ldati']['mixed24']
    out, meta = {}, {}
    for pt, ks, bid in POOLING_CASES:
        ref = rh.run_reference_ldati(vox['mixed24_voxel'], fps=m['fps'], seed=m['seed'], frame_base=m['frame_base'],
                                     bidirectional=bid, pooling_type=pt, pooling_kernel_size=ks)
        name = f"mixed24-{pt}{ks}-{'bi' if bid else 'uni'}"
        out[f'{name}_events_0'] = np.ascontiguousarray(lo.canonicalize(ref[0])).view(np.uint8)
        meta[name] = dict(input='mixed24', fps=m['fps'], seed=m['seed'], frame_base=m['frame_base'], frames=len(ref),
                          pooling_type=pt, pooling_kernel_size=ks, bidirectional=bid,
                          counts=[int(len(r)) for r in ref], sha256=[digest(r) for r in ref])
    np.savez_compressed(os.path.join(HERE, 'ldati_pooling_golden.npz'), **out)
    return meta


def ef_cases():
    out = {}
    meta = {}
    rng = np.random.default_rng(3)
    for name, scale, kp, pct, ceil in (('rgb_small', 0.02, True, 98, 10), ('rgb_mid', 1.0, True, 98, 10),
                                       ('rgb_ceil', 5.0, True, 98, 10), ('gray_mid', 1.0, False, 90, 10),
                                       ('gray_small', 0.02, False, 98, 10)):
        v = (scale * rng.random((4, 2, 10, 30, 38)) ** 2).astype(np.float32)
        fr = rh.run_reference_event_frames(v, 30, ceil, pct, kp)
        out[f'{name}_voxel'] = v
        out[f'{name}_frames'] = fr
        meta[name] = dict(keep_polarity=kp, percentile=pct, ceil=ceil)
    np.savez_compressed(os.path.join(HERE, 'ef_golden.npz'), **out)
    return meta


def unet_cases():
    out = {}
    meta = {}
    V2ce3d = rh.V2ce3d()
    for name, seed, init, shape in (('refinit', 0, 'reference', (1, 4, 2, 36, 44)),
                                    ('lively', 1, 'lively', (2, 16, 2, 20, 28))):
        model = V2ce3d().eval()
        sd = synth.make_state_dict(seed, init)
        assert set(model.state_dict().keys()) == set(sd.keys()), 'state-dict key layout drifted'
        model.load_state_dict(sd)
        g = torch.Generator().manual_seed(100 + seed)
        x = torch.randn(*shape, generator=g)
        with torch.no_grad():
            y0 = model(x).numpy()
            y1 = model(x).numpy()          # second call: spectral-norm state has advanced
        out[f'{name}_x'] = x.numpy()
        out[f'{name}_y0'] = y0
        out[f'{name}_y1'] = y1
        meta[name] = dict(seed=seed, init=init, shape=list(shape))
    np.savez_compressed(os.path.join(HERE, 'unet_golden.npz'), **out)
    return meta


def pipeline_cases():
    """video_to_voxels through the reference driver with a FakeVideoReader."""
    out = {}
    meta = {}
    main = rh.main_module()
    V2ce3d = rh.V2ce3d()
    for name, n_frames, H, W, infer, width, bs in (('center', 20, 28, 40, 'center', 36, 2),
                                                   ('pano', 18, 28, 52, 'pano', 20, 1)):
        model = V2ce3d().eval()
        model.load_state_dict(synth.make_state_dict(2, 'lively'))
        frames = synth.make_video(n_frames, H, W, seed=5)
        with rh.cpu_cuda_shims():
            vox = main.video_to_voxels(model, vidcap=rh.FakeVideoReader(frames), infer_type=infer,
                                       seq_len=16, width=width, height=H, batch_size=bs)
        out[f'{name}_voxel'] = vox.astype(np.float32)
        meta[name] = dict(n_frames=n_frames, H=H, W=W, infer_type=infer, width=width, batch_size=bs,
                          video_seed=5, sd_seed=2, init='lively')
    np.savez_compressed(os.path.join(HERE, 'pipeline_golden.npz'), **out)
    return meta


if __name__ == '__main__':
    import cv2
    if '--only-ldati-options' in sys.argv or '--only-ldati-pooling' in sys.argv:
        # incremental: (re)generate only the LDATI option / pooling goldens, keep the rest of the meta file
        with open(os.path.join(HERE, 'golden_meta.json')) as f:
            meta = json.load(f)
        if '--only-ldati-options' in sys.argv:
            meta['ldati_options'] = ldati_option_cases()
        if '--only-ldati-pooling' in sys.argv:
            meta['ldati_pooling'] = ldati_pooling_cases()
    else:
        meta = dict(versions=dict(torch=torch.__version__, numpy=np.__version__, cv2=cv2.__version__),
                    ldati=ldati_cases(), ef=ef_cases(), unet=unet_cases(), pipeline=pipeline_cases())
        with open(os.path.join(HERE, 'golden_meta.json'), 'w') as f:
            json.dump(meta, f, indent=1)
        meta['ldati_options'] = ldati_option_cases()
        meta['ldati_pooling'] = ldati_pooling_cases()
    with open(os.path.join(HERE, 'golden_meta.json'), 'w') as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta, indent=1)[:3000])
