"""Differential pinning of the oracle against the UNMODIFIED reference, executed live.

Runs only where /root/reference is mounted (the build container; skipped on the GPU box, where the committed
goldens of tests/test_oracle_vs_golden.py stand in).  Randomised small inputs -- shapes, frame rates, value
distributions including zeros, negatives and large counts -- through every option of sample_voxel_statistical that
the oracle restates, and through write_event_frame_video."""
import numpy as np
import pytest

from oracle import ef_oracle, ldati_oracle as lo, ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason='/root/reference is not mounted')


def _random_voxels(rng, kind, B, H, W):
    shape = (B, 2, 10, H, W)
    if kind == 'rand':
        return rng.random(shape, dtype=np.float32)
    if kind == 'sparse':
        return (rng.random(shape, dtype=np.float32) * np.float32(0.05)).astype(np.float32)
    if kind == 'counts':                                   # integer-valued, many multi-event pixel-bins
        return rng.integers(0, 7, shape).astype(np.float32)
    if kind == 'signed':                                   # out-of-contract negatives (the model's output is ReLU'd)
        return (rng.standard_normal(shape) * 1.5).astype(np.float32)
    if kind == 'bursty':                                   # mostly empty with a few large values
        v = np.zeros(shape, np.float32)
        m = rng.random(shape) < 0.08
        v[m] = (rng.random(int(m.sum())) * 12).astype(np.float32)
        return v
    raise ValueError(kind)


CASES = [(kind, strat, bid) for kind in ('rand', 'sparse', 'counts', 'signed', 'bursty')
         for strat in ('slope', 'random', 'none') for bid in (False, True)]


@pytest.mark.parametrize('kind,strategy,bidirectional', CASES)
def test_ldati_oracle_equals_reference_on_random_inputs(kind, strategy, bidirectional):
    seed = CASES.index((kind, strategy, bidirectional)) * 7919 + 13
    rng = np.random.default_rng(seed)
    for trial in range(3):
        B, H, W = int(rng.integers(1, 4)), int(rng.integers(1, 20)), int(rng.integers(1, 24))
        fps = int(rng.choice([24, 25, 30, 50, 60, 120, 240]))
        v = _random_voxels(rng, kind, B, H, W)
        if float(np.max(lo.relocate_counts_bidirectional(v)[0] if bidirectional else lo.relocate_counts(v)[0])) < 1 \
                and strategy != 'none':
            v[0, 0, 3, 0, 0] += np.float32(2.5)            # torch.rand needs a non-empty last dimension to matter
        ref = rh.run_reference_ldati(v, fps=fps, seed=seed + trial, frame_base=trial,
                                     additional_events_strategy=strategy, bidirectional=bidirectional)
        ora = lo.sample_voxel_statistical_oracle(v, fps=fps, seed=seed + trial, frame_base=trial, flavor='cpu',
                                                 additional_events_strategy=strategy, bidirectional=bidirectional)
        assert len(ref) == len(ora) == B
        for i, (r, o) in enumerate(zip(ref, ora)):
            assert lo.events_equal_modulo_ties(r, o), \
                f'{kind} {strategy} bidirectional={bidirectional} trial {trial} shape {(B, H, W)} fps {fps} frame {i}'


@pytest.mark.parametrize('pooling_type,kernel_size', [('weighted', 3), ('avg', 3), ('avg', 5), ('avg', 7)])
@pytest.mark.parametrize('bidirectional', [False, True])
def test_ldati_pooling_oracle_equals_reference_on_random_inputs(pooling_type, kernel_size, bidirectional):
    rng = np.random.default_rng(kernel_size * 31 + int(bidirectional) + len(pooling_type))
    for kind in ('counts', 'bursty', 'rand', 'signed'):
        B, H, W = int(rng.integers(1, 3)), int(rng.integers(1, 16)), int(rng.integers(1, 20))
        fps = int(rng.choice([24, 30, 60, 120]))
        v = _random_voxels(rng, kind, B, H, W)
        v[0, 0, 3, 0, 0] += np.float32(2.5)
        kw = dict(fps=fps, seed=5, frame_base=1, bidirectional=bidirectional, pooling_type=pooling_type,
                  pooling_kernel_size=kernel_size)
        ref = rh.run_reference_ldati(v, **kw)
        ora = lo.sample_voxel_statistical_oracle(v, flavor='cpu', **kw)
        for i, (r, o) in enumerate(zip(ref, ora)):
            assert lo.events_equal_modulo_ties(r, o), f'{pooling_type}{kernel_size} {kind} shape {(B, H, W)} frame {i}'


@pytest.mark.parametrize('keep_polarity', [True, False])
def test_event_frame_oracle_equals_reference_on_random_inputs(keep_polarity):
    rng = np.random.default_rng(5 + int(keep_polarity))
    for trial in range(4):
        N, H, W = int(rng.integers(1, 5)), int(rng.integers(2, 20)), int(rng.integers(2, 24))
        v = _random_voxels(rng, ['rand', 'counts', 'bursty', 'sparse'][trial], N, H, W)
        v = np.maximum(v, 0)
        ceil, pct = int(rng.choice([3, 10, 50])), int(rng.choice([50, 90, 98, 100]))
        ref = rh.run_reference_event_frames(v, 30, ceil, pct, keep_polarity)
        got, _ub, _ = ef_oracle.event_frames_oracle(v, ceil, pct, keep_polarity)
        assert np.array_equal(ref, got), f'trial {trial} shape {(N, H, W)} ceil {ceil} percentile {pct}'


def test_unet_oracle_equals_reference_model_over_three_calls():
    """The folded fp32 restatement vs the reference's own V2ce3d (BatchNorm modules, SpectralNorm wrappers) on a fresh
    random checkpoint and input: three consecutive calls, because the spectral-norm state advances on every forward,
    eval mode included.  Tolerance: fp32 re-association of the folded scale/shift, rel-L2 < 2e-6."""
    import torch
    from oracle import synth
    from oracle.unet_oracle import UNetOracle
    sd = synth.make_state_dict(23, 'lively')
    model = rh.V2ce3d()().eval()
    model.load_state_dict(sd)
    ora = UNetOracle(sd)
    g = torch.Generator().manual_seed(7)
    for call in range(3):
        x = torch.randn(1, 5, 2, 17, 23, generator=g)      # odd sizes: every level of the 17->9->5->3->2 chain is ragged
        with torch.no_grad():
            want = model(x)
        got = ora.forward(x)
        assert got.shape == want.shape == (1, 5, 20, 17, 23)
        rel = float((got - want).norm() / want.norm())
        assert rel < 2e-6, f'call {call}: rel-L2 {rel}'
