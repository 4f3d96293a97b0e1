"""Run the UNMODIFIED reference (mounted read-only at /root/reference) on CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Used only in the build container
(by tests/golden/make_golden.py and by the CPU tests that are skipped when
/root/reference is absent).  Nothing here is copied from the reference: the
modules are imported from where they lie, with the four shims of SURVEY.md F2:

  1. ``pathlib2`` (not installed) is aliased to ``pathlib``;
  2. the reference's namespace package ``scripts`` is loaded under the private
     name ``_v2ce_ref`` so it cannot collide with this repo's own ``scripts``;
  3. on a CPU-only host ``Tensor.cuda()`` / ``Module.to('cuda')`` are identity
     (inside the ``cpu_cuda_shims`` context only);
  4. ``v2ce.py`` reads a module-global ``logger`` that it only creates under
     ``__main__`` (v2ce.py:305-306); we inject it.
"""
import contextlib
import importlib
import importlib.util
import logging
import os
import pathlib
import sys
import types

import numpy as np

REF_ROOT = os.environ.get('V2CE_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'scripts', 'LDATI.py'))


_cache = {}


def _load():
    if _cache:
        return _cache
    if not available():
        raise RuntimeError(f'reference not found under {REF_ROOT}')
    sys.modules.setdefault('pathlib2', pathlib)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k == 'scripts' or k.startswith('scripts.')}
    pkg = types.ModuleType('scripts')
    pkg.__path__ = [os.path.join(REF_ROOT, 'scripts')]
    sys.modules['scripts'] = pkg
    try:
        ldati = importlib.import_module('scripts.LDATI')
        v2ce3d = importlib.import_module('scripts.v2ce_3d')
        sn = importlib.import_module('scripts.spectral_norm')
        importlib.import_module('scripts.video_reader')
        spec = importlib.util.spec_from_file_location('_v2ce_ref_main', os.path.join(REF_ROOT, 'v2ce.py'))
        main = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(main)
        main.logger = logging.getLogger('V2CE')
    finally:
        for k in list(sys.modules):
            if k == 'scripts' or k.startswith('scripts.'):
                sys.modules['_v2ce_ref' + k[len('scripts'):]] = sys.modules.pop(k)
        sys.modules.update(saved)
    _cache.update(ldati=ldati, v2ce3d=v2ce3d, main=main, sn=sn)
    return _cache


def ldati_module():
    return _load()['ldati']


def main_module():
    return _load()['main']


def V2ce3d():
    return _load()['v2ce3d'].V2ce3d


@contextlib.contextmanager
def cpu_cuda_shims():
    import torch
    if torch.cuda.is_available():
        yield
        return
    old_cuda = torch.Tensor.cuda
    old_to = torch.nn.Module.to

    def to(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, str) and x.startswith('cuda')))
        k = {kk: v for kk, v in k.items() if not (kk == 'device' and str(v).startswith('cuda'))}
        return old_to(self, *a, **k) if (a or k) else self

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.to = to
    try:
        yield
    finally:
        torch.Tensor.cuda = old_cuda
        torch.nn.Module.to = old_to


@contextlib.contextmanager
def injected_draws(seed, frame_base=0):
    """Serve the reference's torch.rand((B,2,9,H,W,M)) call (LDATI.py:171) from the
    counter-based stream of oracle/philox.py."""
    import torch
    from . import philox
    old = torch.rand

    def fake_rand(*size, **kw):
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        shape = tuple(int(s) for s in shape)
        if len(shape) != 6:
            return old(*size, **kw)
        B, P, C, H, W, M = shape
        assert P == 2 and C == 9
        d = philox.dense_draws(frame_base, B, H, W, M, seed) if M > 0 else np.zeros(shape, np.float32)
        return torch.from_numpy(d).to(kw.get('device', 'cpu'))

    torch.rand = fake_rand
    try:
        yield
    finally:
        torch.rand = old


def run_reference_ldati(y, fps=30, seed=0, frame_base=0, additional_events_strategy='slope', bidirectional=False,
                        pooling_type='none', pooling_kernel_size=3):
    """sample_voxel_statistical exactly as v2ce.py:356 calls it, on CPU, with injected draws."""
    import torch
    import warnings
    ld = ldati_module()
    yt = torch.as_tensor(np.asarray(y))
    with injected_draws(seed, frame_base), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out = ld.sample_voxel_statistical(yt, fps=fps, bidirectional=bidirectional,
                                          additional_events_strategy=additional_events_strategy,
                                          pooling_type=pooling_type, pooling_kernel_size=pooling_kernel_size)
    return [np.asarray(r) for r in out]


_baseline_cache = {}


def baseline_module():
    """train/scripts/stage2/sample_methods/random_even_sample.py, loaded from where it lies (its h5py / pandas imports
    are not needed by the sampler and may be missing: stubbed for the import only)."""
    if 'baseline' not in _baseline_cache:
        import importlib.machinery
        stubs = {}
        for name in ('h5py', 'pandas'):
            try:
                importlib.import_module(name)
            except Exception:                          # noqa: BLE001
                stubs[name] = sys.modules[name] = types.ModuleType(name)
                sys.modules[name].__spec__ = importlib.machinery.ModuleSpec(name, None)
        try:
            spec = importlib.util.spec_from_file_location(
                '_v2ce_ref_baseline', os.path.join(REF_ROOT, 'train', 'scripts', 'stage2', 'sample_methods', 'random_even_sample.py'))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        finally:
            for name in stubs:
                sys.modules.pop(name, None)
        _baseline_cache['baseline'] = mod
    return _baseline_cache['baseline']


def run_reference_baseline(y, fps=30, seed=0, frame_base=0, even=False, random=False):
    """sample_voxel_baseline on CPU with its torch.rand / torch.bernoulli calls served from the counter-based streams
    of oracle/baseline_oracle.py: rand of a 5-D shape = integer-part uniforms, of a 4-D shape = fractional-part uniforms,
    bernoulli is called once per (frame, bin, plane) in the order frame, bin, negative plane, positive plane."""
    import torch
    import warnings
    from . import philox
    from .baseline_oracle import NB, pixel_bin_index
    mod = baseline_module()
    yt = torch.as_tensor(np.asarray(y))
    B, P, C, H, W = yt.shape
    hw = H * W

    def grid_idx():
        f = np.arange(B).reshape(B, 1, 1, 1) + frame_base
        p = np.arange(2).reshape(1, 2, 1, 1)
        c = np.arange(C).reshape(1, 1, C, 1)
        pix = np.arange(hw).reshape(1, 1, 1, hw)
        return ((f.astype(np.uint64) * np.uint64(2) + p.astype(np.uint64)) * np.uint64(NB) + c.astype(np.uint64)) * np.uint64(hw) + \
            pix.astype(np.uint64)

    old_rand, old_bern = torch.rand, torch.bernoulli
    calls = [0]

    def fake_rand(*size, **kw):
        shape = tuple(int(s) for s in (size[0] if len(size) == 1 and not isinstance(size[0], int) else size))
        idx = grid_idx()
        if len(shape) == 5:                              # (B*P, C, H, W, M)
            M = shape[-1]
            out = np.empty((B, 2, C, hw, M), np.float32)
            for j in range(M):
                out[..., j] = philox.uniform_from_index(idx, np.full(idx.shape, j, np.int64), seed, 0)
            return torch.from_numpy(out.reshape(shape))
        assert len(shape) == 4                           # (B*P, C, H, W)
        return torch.from_numpy(philox.uniform_from_index(idx, np.zeros(idx.shape, np.int64), seed, 1).reshape(shape))

    def fake_bernoulli(prob):
        k = calls[0]
        calls[0] += 1
        b, c, p = k // (2 * C), (k // 2) % C, 1 - (k % 2)
        idx = pixel_bin_index(frame_base + b, p, c, np.arange(hw), hw)
        u = philox.uniform_from_index(idx, np.zeros(hw, np.int64), seed, 2).reshape(H, W)
        return (torch.from_numpy(u) < prob).to(prob.dtype)

    torch.rand, torch.bernoulli = fake_rand, fake_bernoulli
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            out = mod.sample_voxel_baseline(yt, fps=fps, even=even, random=random)
    finally:
        torch.rand, torch.bernoulli = old_rand, old_bern
    return [np.asarray(r) for r in out]


from synth_inputs import FakeVideoReader          # noqa: E402,F401  (in-memory reader; lives with the input generators)


def run_reference_event_frames(voxel, fps=30, ceil=10, percentile=98, keep_polarity=True):
    """write_event_frame_video (v2ce.py:241-280) with cv2.VideoWriter replaced by a
    recorder, so the exact uint8 BGR frames handed to the encoder are captured."""
    import io
    import contextlib as _ctx
    main = main_module()
    frames = []

    class _Recorder:
        def __init__(self, *a, **k):
            pass

        def write(self, f):
            frames.append(np.array(f, copy=True))

        def release(self):
            pass

    cv2 = main.cv2
    old = cv2.VideoWriter
    cv2.VideoWriter = _Recorder
    try:
        with _ctx.redirect_stdout(io.StringIO()):
            main.write_event_frame_video(np.asarray(voxel), '/dev/null', fps, ceil, percentile, keep_polarity)
    finally:
        cv2.VideoWriter = old
    return np.stack(frames)
