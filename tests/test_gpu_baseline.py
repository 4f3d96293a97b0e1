"""GPU parity of the baseline samplers 'random' / 'even'  against the oracle, which
tests/test_baseline_oracle_live.py pins to the unmodified reference
(train/scripts/stage2/sample_methods/random_even_sample.py:118-170).  Bit-exact on all four record fields, row by row."""
import numpy as np
import pytest
import torch

from oracle import baseline_oracle as bo, synth

pytestmark = pytest.mark.gpu


def _run(v, **kw):
    from v2ce_toolbox_b200.sample_methods.random_even_sample import sample_voxel_baseline
    return sample_voxel_baseline(torch.as_tensor(v).cuda(), **kw)


@pytest.mark.parametrize('mode', ['random', 'even'])
@pytest.mark.parametrize('kind,F,H,W,fps', [('rand', 3, 40, 52, 30), ('randint', 2, 24, 30, 24), ('mixed', 2, 33, 47, 120),
                                            ('sparse', 2, 9, 11, 30), ('mixed', 1, 260, 346, 30)])
def test_events_bit_exact_vs_oracle(mode, kind, F, H, W, fps):
    v = synth.make_voxels(kind, F, H, W, seed=17)
    kw = dict(even=mode == 'even', random=mode == 'random')
    got = _run(v, fps=fps, seed=5, frame_base=4, **kw)
    want = bo.sample_voxel_baseline_oracle(v, fps=fps, seed=5, frame_base=4, flavor='cuda', **kw)
    assert len(got) == len(want) == F
    n = 0
    for a, b in zip(got, want):
        assert a.dtype.itemsize == 13 and len(a) == len(b)
        for f in ('timestamp', 'x', 'y', 'polarity'):
            assert np.array_equal(np.asarray(a[f]), np.asarray(b[f])), (mode, kind, f)
        assert (np.diff(a['timestamp']) >= 0).all()
        n += len(a)
    assert n > 0 or kind == 'sparse'


def test_count_conservation_and_ranges():
    """floor(y) integer-part events per pixel-bin at least; every event inside its frame; x, y, polarity in range."""
    v = synth.make_voxels('randint', 2, 32, 48, seed=3)
    got = _run(v, random=True, fps=30, seed=1)
    for f, ev in enumerate(got):
        assert len(ev) == int(np.floor(v[f]).sum())                     # integers: no fractional parts
        assert ev['timestamp'].min() >= 0 and ev['timestamp'].max() <= 33334
        assert ev['x'].max() < 48 and ev['y'].max() < 32 and set(np.unique(ev['polarity'])) <= {0, 1}


def test_rejects_bad_calls():
    from v2ce_toolbox_b200 import V2ceError
    from v2ce_toolbox_b200.sample_methods.random_even_sample import sample_voxel_baseline
    with pytest.raises(AssertionError):
        sample_voxel_baseline(torch.zeros(1, 2, 10, 4, 4, device='cuda'))
    with pytest.raises(V2ceError):
        sample_voxel_baseline(torch.zeros(1, 2, 10, 4, 4), random=True)
    with pytest.raises(V2ceError):                                        # -0.5 with even=True: the reference's timestamp is -inf
        sample_voxel_baseline(torch.full((1, 2, 10, 4, 4), -0.5, device='cuda'), even=True, seed=3)
    assert [len(e) for e in sample_voxel_baseline(torch.zeros(2, 2, 10, 4, 4, device='cuda'), even=True)] == [0, 0]
