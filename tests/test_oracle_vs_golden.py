"""Pin the CPU oracle (oracle/) against fixtures produced by the unmodified
reference (tests/golden/make_golden.py) and the reference's only known-answer
vector (train/scripts/stage2/vis_stage2.ipynb cells 1-2)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import ef_oracle, ldati_oracle as lo, philox, pipeline_oracle, synth
from oracle.unet_oracle import UNetOracle


def test_philox_known_answers():
    for ctr, key, out in philox.KAT:
        r = philox.philox4x32_10(*[np.uint64(c) for c in ctr], key[0], key[1])
        assert tuple(int(x) for x in r) == out


def test_philox_uniform_range():
    idx = np.arange(100000, dtype=np.uint64)
    u = philox.uniform_from_index(idx, idx % np.uint64(7), 42)
    assert u.dtype == np.float32 and u.min() >= 0 and u.max() < 1
    assert abs(float(u.mean()) - 0.5) < 5e-3


def test_notebook_known_answer():
    y = np.zeros((1, 2, 10, 1, 1), np.float32)
    y[0, 0, :, 0, 0] = [0, 0, .9179, .0821, .9962, .0038, .5287, 2.8454, 1.6884, .9375]
    n, _ = lo.relocate_counts(y)
    assert n[0, 0, :, 0, 0].tolist() == [0, 0, 1, 0, 1, 0, 1, 3, 2]
    for flavor in ('cpu', 'cuda'):
        ev = lo.sample_voxel_statistical_oracle(y, fps=30, flavor=flavor)[0]
        assert len(ev) == 8 and (ev['polarity'] == 1).all()
        assert np.sort(ev['timestamp'])[:3].tolist() == [7711, 14828, 23967]


SMALL = ['kat', 'rand', 'randint', 'sparse', 'mixed24', 'mixed120']


@pytest.mark.parametrize('name', SMALL)
def test_ldati_oracle_matches_reference(name, golden, golden_meta):
    g = golden('ldati')
    m = golden_meta['ldati'][name]
    ora = lo.sample_voxel_statistical_oracle(g[f'{name}_voxel'], fps=m['fps'], seed=m['seed'],
                                             frame_base=m['frame_base'], flavor='cpu')
    assert len(ora) == m['frames']
    for i, o in enumerate(ora):
        ref = g[f'{name}_events_{i}'].view(lo.EVENT_DTYPE)
        assert len(o) == m['counts'][i]
        assert np.array_equal(lo.canonicalize(o), ref), f'{name} frame {i}'
        assert o.dtype.itemsize == 13


OPTION_CASES = [f"{inp}-{strat}-{d}" for inp in ('mixed24', 'randint') for strat in ('slope', 'random', 'none')
                for d in ('uni', 'bi') if (strat, d) != ('slope', 'uni')]


@pytest.mark.parametrize('name', OPTION_CASES)
def test_ldati_oracle_options_match_reference(name, golden, golden_meta):
    """additional_events_strategy 'random' / 'none' and bidirectional=True against the
    unmodified reference's events."""
    m = golden_meta['ldati_options'][name]
    g = golden('ldati_options')
    v = golden('ldati')[f"{m['input']}_voxel"]
    ora = lo.sample_voxel_statistical_oracle(v, fps=m['fps'], seed=m['seed'], frame_base=m['frame_base'], flavor='cpu',
                                             additional_events_strategy=m['additional_events_strategy'],
                                             bidirectional=m['bidirectional'])
    assert [len(o) for o in ora] == m['counts']
    for i, o in enumerate(ora):
        assert np.array_equal(lo.canonicalize(o), g[f'{name}_events_{i}'].view(lo.EVENT_DTYPE)), f'{name} frame {i}'


@pytest.mark.parametrize('name', ['mixed24-weighted3-uni', 'mixed24-avg3-uni', 'mixed24-avg5-bi'])
def test_ldati_oracle_pooling_matches_reference(name, golden, golden_meta):
    """pooling_type 'weighted' / 'avg' (LDATI.py:176-183): the slope is fitted on spatially pooled counts."""
    m = golden_meta['ldati_pooling'][name]
    v = golden('ldati')['mixed24_voxel']
    ora = lo.sample_voxel_statistical_oracle(v, fps=m['fps'], seed=m['seed'], frame_base=m['frame_base'], flavor='cpu',
                                             bidirectional=m['bidirectional'], pooling_type=m['pooling_type'],
                                             pooling_kernel_size=m['pooling_kernel_size'])
    assert [len(o) for o in ora] == m['counts']
    assert np.array_equal(lo.canonicalize(ora[0]), golden('ldati_pooling')[f'{name}_events_0'].view(lo.EVENT_DTYPE))
    for o, d in zip(ora, m['sha256']):
        assert hashlib.sha256(np.ascontiguousarray(lo.canonicalize(o)).tobytes()).hexdigest() == d


def test_pool_counts_definitions():
    rng = np.random.default_rng(0)
    n = rng.integers(0, 9, (2, 9, 6, 7))
    w = lo.pool_counts(n, 'weighted')
    assert w[0, 0, 0, 0] == np.float32((4 * n[0, 0, 0, 0] + 2 * n[0, 0, 0, 1] + 2 * n[0, 0, 1, 0] + n[0, 0, 1, 1]) / 16)
    a = lo.pool_counts(n, 'avg', 3)
    assert a[1, 4, 2, 3] == np.float32(n[1, 4, 1:4, 2:5].sum()) / np.float32(9)
    assert a[1, 4, 0, 0] == np.float32(n[1, 4, 0:2, 0:2].sum()) / np.float32(9)      # zero padding counts in the divisor
    assert np.array_equal(lo.pool_counts(n, 'avg', 1), n.astype(np.float32))
    assert np.array_equal(lo.pool_counts(n, 'none'), n.astype(np.float32))


def test_bidirectional_relocation_shape():
    """LDATI.py:107-122: bin 4 is never written; bin 8's tendency is the tenth voxel bin itself."""
    v = synth.make_voxels('mixed', 1, 16, 20, seed=2)
    n, tend = lo.relocate_counts_bidirectional(v)
    assert (n[:, :, 4] == 0).all() and (tend[:, :, 4] == 0).all()
    assert np.array_equal(tend[:, :, 8], v[:, :, 9])
    n_uni, tend_uni = lo.relocate_counts(v)
    assert np.array_equal(n[:, :, :4], n_uni[:, :, :4]) and np.array_equal(tend[:, :, :4], tend_uni[:, :, :4])


@pytest.mark.parametrize('kind', ['rand', 'sparse', 'randint'])
def test_ldati_oracle_full_size_digest(kind, golden_meta):
    m = golden_meta['ldati'][f'full_{kind}']
    v = synth.make_voxels(kind, 1, 260, 346, seed=m['voxel_seed'])
    ora = lo.sample_voxel_statistical_oracle(v, fps=m['fps'], seed=m['seed'], frame_base=m['frame_base'],
                                             flavor='cpu')
    assert [len(o) for o in ora] == m['counts']
    for o, d in zip(ora, m['sha256']):
        c = np.ascontiguousarray(lo.canonicalize(o))
        assert hashlib.sha256(c.tobytes()).hexdigest() == d


def test_ldati_flavours_differ_only_rarely():
    v = synth.make_voxels('mixed', 1, 64, 80, seed=3)
    a = lo.sample_voxel_statistical_oracle(v, flavor='cpu', seed=1)[0]
    b = lo.sample_voxel_statistical_oracle(v, flavor='cuda', seed=1)[0]
    assert len(a) == len(b)
    d = np.abs(np.sort(a['timestamp']) - np.sort(b['timestamp']))
    assert d.max() <= 1 and (d != 0).mean() < 1e-2


def test_ldati_properties():
    v = synth.make_voxels('randint', 1, 16, 20, seed=5)
    ev, counts = lo.sample_voxel_statistical_oracle(v, return_seg_counts=True)
    n, _ = lo.relocate_counts(v)
    assert np.clip(n, 0, None).sum() == len(ev[0]) == counts.sum()
    e = ev[0]
    assert e['x'].min() >= 0 and e['x'].max() < 20 and e['y'].min() >= 0 and e['y'].max() < 16
    assert set(np.unique(e['polarity'])) <= {0, 1}
    start = 0
    for c in range(9):                      # time-sorted inside every bin segment
        seg = e['timestamp'][start:start + counts[0, c]]
        assert (np.diff(seg) >= 0).all()
        start += counts[0, c]


@pytest.mark.parametrize('name', ['rgb_small', 'rgb_mid', 'rgb_ceil', 'gray_mid', 'gray_small'])
def test_event_frame_oracle_matches_reference(name, golden, golden_meta):
    g = golden('ef')
    m = golden_meta['ef'][name]
    frames, ub, _ = ef_oracle.event_frames_oracle(g[f'{name}_voxel'], m['ceil'], m['percentile'], m['keep_polarity'])
    assert np.array_equal(frames, g[f'{name}_frames'])


def test_percentile_matches_numpy():
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 1000, 4097):
        v = rng.random(n).astype(np.float32)
        for q in (0, 37, 50, 98, 100):
            a, b, t, _ = ef_oracle.order_statistic_pair(v, q)
            assert ef_oracle.lerp_percentile(a, b, t) == np.percentile(v.astype(np.float64), q)


@pytest.mark.parametrize('name', ['refinit', 'lively'])
def test_unet_oracle_matches_reference(name, golden, golden_meta):
    g = golden('unet')
    m = golden_meta['unet'][name]
    orc = UNetOracle(synth.make_state_dict(m['seed'], m['init']))
    x = torch.from_numpy(g[f'{name}_x'])
    for call in (0, 1):
        y = orc.forward(x).numpy()
        ref = g[f'{name}_y{call}']
        rel = np.linalg.norm(y - ref) / np.linalg.norm(ref)
        assert rel < 2e-6, (name, call, rel)     # fp32 conv; BN folded into scale/shift
    assert np.linalg.norm(g[f'{name}_y0'] - g[f'{name}_y1']) > 0   # SN state really advances


@pytest.mark.parametrize('name', ['center', 'pano'])
def test_pipeline_oracle_matches_reference(name, golden, golden_meta):
    g = golden('pipeline')
    m = golden_meta['pipeline'][name]
    orc = UNetOracle(synth.make_state_dict(m['sd_seed'], m['init']))
    frames = synth.make_video(m['n_frames'], m['H'], m['W'], seed=m['video_seed'])
    vox = pipeline_oracle.video_to_voxels(orc.forward, frames, m['infer_type'], 16, m['width'], m['H'],
                                          m['batch_size'])
    ref = g[f'{name}_voxel']
    assert vox.shape == ref.shape
    assert np.linalg.norm(vox - ref) / np.linalg.norm(ref) < 2e-6


def test_window_schedule():
    starts, mode = pipeline_oracle.window_starts(321)
    assert len(starts) == 20 and mode == 0 and starts[-1] == 304
    starts, mode = pipeline_oracle.window_starts(600)
    assert len(starts) == 38 and mode == 7 and starts[-1] == 37 * 16 - 9
    assert pipeline_oracle.pano_tiles(462) == [(0, 346, 346), (116, 462, 116)]
    assert pipeline_oracle.frame_offset_us(1, 30) == 33333 and pipeline_oracle.frame_offset_us(3, 30) == 100000
