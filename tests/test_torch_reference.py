"""torch_reference.py (the reference's op sequence through torch's own kernels) is itself pinned on CPU: against
the oracle's torch-CPU flavour (which tests/test_oracle_vs_golden.py pins to the unmodified reference) and, where
/root/reference is mounted, against the unmodified reference live.  On the GPU box the same functions run on
torch-CUDA and are what the `cuda` flavour of the kernels is compared with end to end
(tests/test_gpu_torch_reference.py)."""
import numpy as np
import pytest
import torch

import torch_reference as tr
from oracle import ldati_oracle as lo, ref_harness, synth
from oracle.unet_oracle import UNetOracle


def _vox(kind, B, H, W, seed):
    rng = np.random.default_rng(seed)
    if kind == 'rand':
        return rng.random((B, 2, 10, H, W), dtype=np.float32)
    if kind == 'randint':
        return rng.integers(0, 10, (B, 2, 10, H, W)).astype(np.float32)
    v = rng.random((B, 2, 10, H, W), dtype=np.float32) * 3
    v[rng.random(v.shape) < 0.5] = 0
    return v.astype(np.float32)


@pytest.mark.parametrize('kind', ['rand', 'randint', 'mixed'])
@pytest.mark.parametrize('opts', [dict(), dict(additional_events_strategy='random'), dict(additional_events_strategy='none'),
                                  dict(bidirectional=True), dict(pooling_type='weighted'),
                                  dict(pooling_type='avg', pooling_kernel_size=5)])
def test_ldati_torch_cpu_equals_cpu_oracle(kind, opts):
    B, H, W = 2, 9, 11
    vox = _vox(kind, B, H, W, 3)
    draws = np.random.default_rng(5).random((B, 2, 9, H, W, 24), dtype=np.float32)
    got = tr.sample_voxel_statistical_torch(torch.from_numpy(vox), fps=30, draws=torch.from_numpy(draws), **opts)
    want = lo.sample_voxel_statistical_oracle(vox, fps=30, flavor='cpu', draws=draws, **opts)
    for a, b in zip(got, want):
        assert len(a) == len(b)
        # torch-CPU's vectorised float32 sqrt is not correctly rounded; the oracle's cpu flavour
        # calls torch for it, so everything matches row for row
        for f in ('timestamp', 'x', 'y', 'polarity'):
            assert np.array_equal(np.asarray(a[f]), np.asarray(b[f])), f


@pytest.mark.skipif(not ref_harness.available(), reason='/root/reference is not mounted')
@pytest.mark.parametrize('fps', [30, 60, 240])
def test_ldati_torch_cpu_equals_reference_live(fps):
    B, H, W = 2, 8, 12
    vox = _vox('mixed', B, H, W, fps)
    ref = ref_harness.run_reference_ldati(vox, fps=fps, seed=7)
    from oracle import philox
    draws = philox.dense_draws(0, B, H, W, 40, 7)
    got = tr.sample_voxel_statistical_torch(torch.from_numpy(vox), fps=fps, draws=torch.from_numpy(draws))
    for a, b in zip(got, ref):
        assert lo.events_equal_modulo_ties(a, b)


def test_unet_torch_cpu_equals_oracle():
    sd = synth.make_state_dict(3, 'lively')
    x = torch.randn(1, 4, 2, 20, 24, generator=torch.Generator().manual_seed(1))
    a, b = tr.TorchV2ce3d(sd), UNetOracle(sd)
    for _ in range(2):                                       # two calls: the spectral-norm state advances
        ya, yb = a(x), b.forward(x)
        assert float((ya - yb).norm() / yb.norm()) < 2e-6


def test_event_frames_torch_cpu_close_to_oracle():
    from oracle import ef_oracle
    vox = _vox('rand', 3, 10, 12, 1)
    frames, ub = tr.event_frames_torch(torch.from_numpy(vox), 10, 98, True)
    want, want_ub, _ = ef_oracle.event_frames_oracle(vox, 10, 98, True)
    assert abs(ub - want_ub) < 1e-5 * want_ub
    assert np.abs(frames.numpy().astype(int) - want.astype(int)).max() <= 1
