"""The baseline samplers' oracle (oracle/baseline_oracle.py) against the UNMODIFIED reference
(train/scripts/stage2/sample_methods/random_even_sample.py:118-170) run live on CPU with the same injected draws;
skipped where /root/reference is not mounted.  Equal up to the reference's undefined order of equal timestamps."""
import numpy as np
import pytest

from oracle import baseline_oracle as bo, ldati_oracle as lo, ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason='/root/reference is not mounted')


def _vox(kind, B, H, W, seed):
    rng = np.random.default_rng(seed)
    if kind == 'rand':
        return rng.random((B, 2, 10, H, W), dtype=np.float32)
    if kind == 'mixed':
        v = rng.random((B, 2, 10, H, W), dtype=np.float32) * 4
        v[rng.random(v.shape) < 0.4] = 0
        return v.astype(np.float32)
    return (rng.random((B, 2, 10, H, W), dtype=np.float32) * 3 - 0.5).astype(np.float32)      # with negative values


@pytest.mark.parametrize('mode', ['random', 'even'])
@pytest.mark.parametrize('kind,fps', [('rand', 30), ('mixed', 30), ('mixed', 120), ('signed', 24)])
def test_oracle_equals_reference(mode, kind, fps):
    v = _vox(kind, 2, 7, 9, fps)
    kw = dict(even=mode == 'even', random=mode == 'random')
    ref = ref_harness.run_reference_baseline(v, fps=fps, seed=11, frame_base=3, **kw)
    got = bo.sample_voxel_baseline_oracle(v, fps=fps, seed=11, frame_base=3, flavor='cpu', **kw)
    assert len(ref) == len(got) == 2
    for a, b in zip(got, ref):
        assert len(a) == len(b) and len(a) > 0
        assert lo.events_equal_modulo_ties(a, b)
