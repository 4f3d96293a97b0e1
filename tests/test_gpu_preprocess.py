"""GPU parity of the device-side image pre-processing (csrc/preproc.cu, v2ce.py:45-64 for frames that need resizing)
against the oracle (pinned to OpenCV / the reference by tests/test_resize_oracle.py) and the host path.

First hardware run: round-1 driver box, 9/9 green; the device route is stream_clip's default for frames at another
resolution since round 2."""
import numpy as np
import pytest
import torch

from oracle import resize_oracle as ro, synth
from oracle.ref_harness import FakeVideoReader

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize('shape,height', [((2, 5, 72, 128), 26), ((1, 17, 54, 96), 260), ((1, 3, 1080, 1920), 260),
                                          ((1, 3, 260, 346), 260), ((3, 2, 37, 53), 26), ((1, 4, 300, 301), 260)])
def test_image_units_bit_exact(shape, height):
    """/255, cv2-exact bilinear resize (down, up, identity, ragged), pair stacking, Normalize: every float identical."""
    from v2ce_toolbox_b200 import v2ce as drv
    from v2ce_toolbox_b200.preprocess import image_units_device
    rng = np.random.default_rng(shape[2] + height)
    frames = rng.integers(0, 256, shape, dtype=np.uint8)
    got = image_units_device(torch.from_numpy(frames).cuda(), height).cpu().numpy()
    for w in range(shape[0]):
        want = ro.image_units(frames[w], height)
        assert got[w].shape == want.shape
        assert np.array_equal(got[w].view(np.uint32), want.view(np.uint32)), f'window {w}: {(got[w] != want).sum()} differ'
        host = drv.image_pre_processing(frames[w], height).numpy()          # the driver's cv2 path
        assert np.array_equal(got[w].view(np.uint32), host.view(np.uint32))


def test_image_units_rejects_bad_inputs():
    from v2ce_toolbox_b200 import V2ceError
    from v2ce_toolbox_b200.preprocess import image_units_device
    with pytest.raises(V2ceError):
        image_units_device(torch.zeros(1, 3, 8, 8, dtype=torch.uint8), 8)              # CPU tensor
    with pytest.raises(V2ceError):
        image_units_device(torch.zeros(1, 3, 8, 8, device='cuda'), 8)                  # not uint8
    with pytest.raises(V2ceError):
        image_units_device(torch.zeros(1, 1, 8, 8, dtype=torch.uint8, device='cuda'), 8)   # a window needs two frames


@pytest.mark.parametrize('infer_type,width', [('center', 36), ('pano', 20)])
def test_stream_clip_device_resize_equals_host_resize(infer_type, width):
    """A clip at another resolution: uploading raw frames and resizing on the device yields the same event stream and
    preview frames, byte for byte, as the default host cv2 path."""
    from v2ce_toolbox_b200 import v2ce as drv
    from v2ce_toolbox_b200.scripts.v2ce_3d import V2ce3d
    H = 28
    frames = synth.make_video(20, 60, 112, seed=4)             # 60x112 -> 28x52
    out = []
    for dev_resize in (False, True):
        m = V2ce3d()
        m.load_state_dict(synth.make_state_dict(6, 'lively'))
        m = m.eval().to('cuda')
        out.append(drv.stream_clip(m, vidcap=FakeVideoReader(frames), infer_type=infer_type, seq_len=16, width=width,
                                   height=H, batch_size=2, fps=30, seed=3, device_resize=dev_resize))
    a, b = out
    assert a.n_pairs == b.n_pairs == 19 and len(a.event_stream) == len(b.event_stream) > 0
    assert np.array_equal(a.event_stream.view(np.uint8), b.event_stream.view(np.uint8))
    assert a.ef_upper_bound == b.ef_upper_bound and np.array_equal(a.ef_frames, b.ef_frames)
