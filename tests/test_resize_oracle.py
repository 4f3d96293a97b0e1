"""Pin oracle/resize_oracle.py (cv2's float32 bilinear resize restated, and v2ce.py:45-64 on top of it) against the
installed OpenCV -- a library the reference calls, present here and on the GPU box -- and, where /root/reference is
mounted, against the reference's own image_pre_processing."""
import numpy as np
import pytest

from oracle import ref_harness as rh, resize_oracle as ro

SHAPES = [((1080, 1920), (260, 462)), ((720, 1280), (260, 462)), ((480, 640), (260, 346)), ((260, 346), (260, 346)),
          ((100, 133), (260, 345)), ((37, 53), (26, 37)), ((8, 16), (8, 40)), ((8, 16), (20, 16)), ((2, 5), (3, 7)),
          ((300, 301), (260, 260)), ((1080, 1920), (1080, 1920)), ((64, 48), (17, 200))]


@pytest.mark.parametrize('src,dst', SHAPES)
def test_resize_oracle_equals_cv2(src, dst):
    import cv2
    rng = np.random.default_rng(src[0] * 7 + dst[1])
    for kind in ('uniform', 'u8'):
        img = rng.random(src, dtype=np.float32) if kind == 'uniform' else \
            rng.integers(0, 256, src).astype(np.float32) / np.float32(255)
        want = cv2.resize(img, (dst[1], dst[0]))
        got = ro.cv2_resize_linear_f32(img, dst[1], dst[0])
        assert got.dtype == np.float32 and got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f'{kind}: {(got != want).sum()} of {got.size} differ'


def test_identity_when_the_height_already_matches():
    """a 346x260 input goes through cv2.resize unchanged, which is what lets the native path skip it."""
    rng = np.random.default_rng(0)
    img = rng.random((260, 346), dtype=np.float32)
    assert np.array_equal(ro.cv2_resize_linear_f32(img, 346, 260), img)


@pytest.mark.parametrize('shape,height', [((5, 72, 128), 26), ((3, 260, 346), 260), ((4, 54, 96), 260)])
def test_image_units_equal_the_driver_and_the_reference(shape, height):
    from v2ce_toolbox_b200 import v2ce as drv
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, shape, dtype=np.uint8)
    got = ro.image_units(frames, height)
    ours = drv.image_pre_processing(frames, height).numpy()            # the driver's host path (cv2)
    assert got.shape == ours.shape and np.array_equal(got.view(np.uint32), ours.view(np.uint32))
    if rh.available():
        ref = rh.main_module().image_pre_processing(frames, height).numpy()
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
