"""GPU parity of LDATI's pooling_type 'weighted' / 'avg' (LDATI.py:176-183: the slope of a multi-event pixel-bin is
fitted on spatially pooled counts) against the oracle and the reference goldens.

The oracle is pinned to the reference (goldens + live differential tests).  First hardware run: round-1 driver box,
16/16 green; the route is on by default since round 2.

Flavour note: in the torch-CUDA flavour the reference pools with cuDNN.  'weighted' is exact in any arithmetic (dyadic
weights on small integers); for 'avg' the oracle assumes float32 (not TF32) arithmetic in the slope's conv1d --
tests/test_gpu_torch_semantics.py::test_pooled_slope_ops_are_float32_on_cuda checks exactly that against torch-CUDA
(green on the B200: torch runs those ops in float32)."""
import numpy as np
import pytest
import torch

from oracle import ldati_oracle as lo, synth

pytestmark = [pytest.mark.gpu]


def _run(vox, **kw):
    from v2ce_toolbox_b200.scripts.LDATI import sample_voxel_statistical
    return sample_voxel_statistical(torch.as_tensor(vox).cuda(), **kw)



@pytest.mark.parametrize('pooling_type,kernel_size', [('weighted', 3), ('avg', 3), ('avg', 5)])
@pytest.mark.parametrize('kind,F,H,W,bidirectional', [('mixed', 2, 33, 47, False), ('randint', 2, 24, 30, False),
                                                      ('mixed', 2, 40, 52, True), ('mixed', 1, 260, 346, False)])
def test_pooling_bit_exact_vs_oracle(pooling_type, kernel_size, kind, F, H, W, bidirectional):
    v = synth.make_voxels(kind, F, H, W, seed=61)
    kw = dict(fps=30, seed=3, frame_base=4, bidirectional=bidirectional, pooling_type=pooling_type,
              pooling_kernel_size=kernel_size)
    got = _run(v, **kw)
    want = lo.sample_voxel_statistical_oracle(v, flavor='cuda', **kw)
    for i, (g, w) in enumerate(zip(got, want)):
        assert len(g) == len(w)
        for f in ('timestamp', 'x', 'y', 'polarity'):
            assert np.array_equal(np.asarray(g[f]), np.asarray(w[f])), f'{pooling_type}{kernel_size} {kind} frame {i} {f}'


@pytest.mark.parametrize('name', ['mixed24-weighted3-uni', 'mixed24-avg3-uni', 'mixed24-avg5-bi'])
def test_pooling_cpu_flavour_against_reference_goldens(name, golden, golden_meta):
    """Kernel in torch-CPU scalar semantics vs the unmodified reference's events (criterion of
    test_gpu_ldati.py::test_cpu_flavour_against_reference_goldens: equal counts, timestamps within 1 us on at most
    1e-3 of the events because torch-CPU's float32 sqrt is not correctly rounded)."""
    m = golden_meta['ldati_pooling'][name]
    v = golden('ldati')['mixed24_voxel']
    got = _run(v, fps=m['fps'], seed=m['seed'], frame_base=m['frame_base'], flavor='cpu',
               bidirectional=m['bidirectional'], pooling_type=m['pooling_type'],
               pooling_kernel_size=m['pooling_kernel_size'])
    assert [len(e) for e in got] == m['counts']
    ref = golden('ldati_pooling')[f'{name}_events_0'].view(lo.EVENT_DTYPE)
    d = np.abs(np.sort(got[0]['timestamp']) - np.sort(ref['timestamp']))
    assert d.max() <= 1 and (d != 0).mean() <= 1e-3


def test_pooling_needs_an_odd_kernel():
    from v2ce_toolbox_b200.scripts.LDATI import sample_voxel_statistical
    y = torch.zeros(1, 2, 10, 4, 4, device='cuda')
    with pytest.raises(ValueError):
        sample_voxel_statistical(y, pooling_type='avg', pooling_kernel_size=4)
    assert [len(e) for e in sample_voxel_statistical(y, pooling_type='weighted')] == [0]
