"""Image pre-processing (v2ce.py:45-64) -- CPU restatement including cv2's float32 bilinear resize.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference resizes every frame with
``cv2.resize(img_f32, (int(W/H*height), height))`` (v2ce.py:59; INTER_LINEAR).  OpenCV is a third-party
dependency of the reference (opencv_contrib_python 4.8.0.76 pinned in train/requirements.txt:11, 4.13.0 installed
here), so its algorithm is restated from its observable behaviour and pinned against the installed library by
tests/test_resize_oracle.py, on the shapes the CLI meets (1080x1920 -> 260x462, 720x1280 -> 260x462,
480x640 -> 260x346, identity) and on ragged small ones:

  coordinates   f = (d + 0.5) * (src / dst) - 0.5 in double; s = floor(f); a = float32(f - s);
                s < 0 -> (0, a = 0); s >= src - 1 -> (src - 1, a = 0); the second tap is min(s + 1, src - 1)
  horizontal    h[y][dx] = fma(S[y][x1] - S[y][x0], ax[dx], S[y][x0])        float32, ONE rounding for the fma
  vertical      out[dy][dx] = fma(h[y1][dx] - h[y0][dx], ay[dy], h[y0][dx])

(the fused multiply-add is what OpenCV's AVX2 / AVX-512 dispatch of the float path executes; with one source row
OpenCV takes a different scalar path, which no video frame meets.)

  cv2_resize_linear_f32   <- cv2.resize(float32 image, (dw, dh)) as called at /root/reference/v2ce.py:59
  image_units             <- /root/reference/v2ce.py:45-64 (image_pre_processing): /255, resize, pair stacking,
                             Normalize(0.153, 0.165) == (x - 0.153) / 0.165 in float32
"""
import numpy as np

F32 = np.float32


def taps(dst, src):
    """(first tap index int64 [dst], second tap index, float32 weight of the second tap)."""
    scale = src / dst                                      # Python double, like cv::resize's scale_x / scale_y
    f = (np.arange(dst, dtype=np.float64) + 0.5) * scale - 0.5
    s = np.floor(f).astype(np.int64)
    a = f - s
    low, high = s < 0, s >= src - 1
    s[low], a[low] = 0, 0.0
    s[high], a[high] = src - 1, 0.0
    return s, np.minimum(s + 1, src - 1), a.astype(F32)


def _fma32(a, b, c):
    """float32 fma(a, b, c): the product of two float32 is exact in float64; the float64 sum is rounded once more to
    float32.  (Double rounding could differ from a true fma only when the float64 sum lands exactly on a float32
    tie, which needs a 29-bit cancellation pattern; the tests against cv2 would expose it.)"""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


def cv2_resize_linear_f32(img, dw, dh):
    """img (H, W) float32 -> (dh, dw) float32, bit-identical to cv2.resize(img, (dw, dh)) for H >= 2."""
    img = np.asarray(img, dtype=F32)
    sh, sw = img.shape
    x0, x1, ax = taps(dw, sw)
    y0, y1, ay = taps(dh, sh)
    h = _fma32((img[:, x1] - img[:, x0]).astype(F32), np.broadcast_to(ax, (sh, dw)), img[:, x0])
    return _fma32((h[y1] - h[y0]).astype(F32), np.broadcast_to(ay[:, None], (dh, dw)), h[y0])


def resized_width(src_h, src_w, height):
    return int(src_w / src_h * height)                     # v2ce.py:59


def image_units(frames_u8, height=260):
    """frames (N, H, W) uint8 -> image units (N-1, 2, height, W') float32 (v2ce.py:45-64)."""
    frames_u8 = np.asarray(frames_u8)
    n, sh, sw = frames_u8.shape
    dw = resized_width(sh, sw, height)
    x = frames_u8.astype(F32) / F32(255)
    r = np.stack([cv2_resize_linear_f32(f, dw, height) for f in x], axis=0)
    units = np.stack([r[:-1], r[1:]], axis=1)
    return ((units - F32(0.153)) / F32(0.165)).astype(F32)
