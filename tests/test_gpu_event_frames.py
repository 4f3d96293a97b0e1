"""GPU parity: event-frame kernels vs the CPU oracle and the reference goldens (bit-exact uint8)."""
import numpy as np
import pytest
import torch

from oracle import ef_oracle, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['rgb_small', 'rgb_mid', 'rgb_ceil', 'gray_mid', 'gray_small'])
def test_frames_match_reference_goldens(name, golden, golden_meta):
    from v2ce_toolbox_b200 import event_frames as ef
    g = golden('ef')
    m = golden_meta['ef'][name]
    frames, ub = ef.event_frames(torch.from_numpy(g[f'{name}_voxel']).cuda(), m['ceil'], m['percentile'],
                                 m['keep_polarity'])
    assert np.array_equal(frames.cpu().numpy(), g[f'{name}_frames'])


@pytest.mark.parametrize('keep', [True, False])
@pytest.mark.parametrize('shape', [(3, 33, 47), (4, 260, 346)])
def test_stages_bit_exact_vs_oracle(keep, shape):
    from v2ce_toolbox_b200 import event_frames as ef
    N, H, W = shape
    v = synth.make_voxels('mixed', N, H, W, seed=17) * np.float32(0.3)
    v[0, :, :, :3] = 0                                  # some exact zeros / non-positive sums
    vd = torch.from_numpy(v).cuda()
    sums = ef.accumulate(vd, keep)
    want_sums = ef_oracle.accumulate(v, keep)
    assert np.array_equal(sums.cpu().numpy().view(np.uint32), want_sums.view(np.uint32))
    for pct in (0, 50, 98, 100):
        ub = ef.upper_bound(sums, pct, 10, keep)
        assert ub == ef_oracle.upper_bound(want_sums, pct, 10, keep), pct
    ub = ef_oracle.upper_bound(want_sums, 98, 10, keep)
    fr = ef.normalize(sums, ub, keep).cpu().numpy()
    assert np.array_equal(fr, ef_oracle.to_bgr_u8(want_sums, ub, keep))


def test_dropin_writer(tmp_path):
    from v2ce_toolbox_b200.event_frames import write_event_frame_video
    v = synth.make_voxels('rand', 3, 64, 80, seed=1)
    p = str(tmp_path / 'ef.mp4')
    frames = write_event_frame_video(v, p, 30, 10, 98, True)
    want, _, _ = ef_oracle.event_frames_oracle(v, 10, 98, True)
    assert np.array_equal(frames, want)
    import os
    assert os.path.getsize(p) > 0


def test_all_zero_raises_like_numpy():
    from v2ce_toolbox_b200 import event_frames as ef
    with pytest.raises(ValueError):
        ef.event_frames(torch.zeros(1, 2, 10, 8, 8, device='cuda'))


def test_count_pass_writes_the_same_sums_as_accumulate():
    """v2ce_ldati_count_ef: the per-polarity event-frame sums written by the LDATI count pass (one read of the voxels
    for both stages, runner.BatchRunner) are bit for bit those of v2ce_ef_accumulate -- V=4 and V=1 pixel paths."""
    from v2ce_toolbox_b200 import event_frames as ef, ldati
    for (n, H, W) in ((3, 20, 28), (2, 9, 11), (2, 260, 346)):
        g = torch.Generator(device='cuda').manual_seed(H)
        vox = torch.rand((n, 2, 10, H, W), generator=g, device='cuda') * 0.3
        eng = ldati.LdatiEngine('cuda')
        params = ldati.make_params(n, H, W, fps=30, seed=1, device='cuda')
        sums = torch.full((n, 2, H, W), float('nan'), device='cuda')
        seg = eng.count(vox, params, ef_sums=sums)
        assert torch.equal(sums, ef.accumulate(vox, True))
        assert torch.equal(seg, ldati.LdatiEngine('cuda').count(vox, params))
