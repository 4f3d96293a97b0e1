"""Host-side I/O around the device path: frame prefetch and the .npz event sink.  CPU only."""
import os
import zipfile
import zlib

import numpy as np
import pytest

from v2ce_toolbox_b200 import sink
from v2ce_toolbox_b200 import v2ce as drv
from v2ce_toolbox_b200.ldati import EVENT_DTYPE


def test_crc32_combine_matches_zlib():
    rng = np.random.default_rng(0)
    for la, lb in ((0, 5), (7, 0), (1, 1), (1000, 12345), (1 << 16, (1 << 18) + 3)):
        a = rng.integers(0, 256, la, dtype=np.uint8).tobytes()
        b = rng.integers(0, 256, lb, dtype=np.uint8).tobytes()
        assert sink.crc32_combine(zlib.crc32(a), zlib.crc32(b), lb) == zlib.crc32(a + b)


def test_save_npz_reads_back_like_np_savez(tmp_path):
    """v2ce.py:371-372: the archive must be what np.savez(path, event_stream=...) writes -- same member name, bytes,
    CRC and zip64 layout -- so np.load and any zip tool read it."""
    rng = np.random.default_rng(1)
    n = 200_003
    ev = np.zeros(n, EVENT_DTYPE)
    ev['timestamp'] = np.sort(rng.integers(0, 1 << 40, n))
    ev['x'], ev['y'], ev['polarity'] = rng.integers(0, 346, n), rng.integers(0, 260, n), rng.integers(0, 2, n)
    p = sink.save_npz(tmp_path / 'clip-events', chunk_bytes=1 << 18, workers=3, event_stream=ev,
                      empty=np.empty(0, EVENT_DTYPE), grid=np.arange(12, dtype=np.float32).reshape(3, 4))
    assert p.endswith('clip-events.npz')
    np.savez(tmp_path / 'ref.npz', event_stream=ev)
    ours, ref = zipfile.ZipFile(p), zipfile.ZipFile(tmp_path / 'ref.npz')
    assert ours.testzip() is None
    a, b = ours.getinfo('event_stream.npy'), ref.getinfo('event_stream.npy')
    assert (a.CRC, a.file_size, a.compress_type, a.flag_bits, a.extract_version) == \
           (b.CRC, b.file_size, b.compress_type, b.flag_bits, b.extract_version)
    assert ours.read('event_stream.npy') == ref.read('event_stream.npy')
    d = np.load(p)
    assert d['event_stream'].dtype == EVENT_DTYPE and np.array_equal(d['event_stream'], ev)
    assert len(d['empty']) == 0 and d['empty'].dtype == EVENT_DTYPE
    assert np.array_equal(d['grid'], np.arange(12, dtype=np.float32).reshape(3, 4))


def _write_pngs(folder, n, h=20, w=30, seed=0):
    import cv2
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    paths = []
    for i, f in enumerate(frames):
        paths.append(os.path.join(folder, f'{i:05d}.png'))
        cv2.imwrite(paths[-1], f)
    return frames, paths


@pytest.mark.parametrize('n_frames', [17, 33, 40, 49])
def test_frame_prefetcher_equals_serial_reader(tmp_path, n_frames):
    """The threaded decode hands out exactly the windows of the serial cv2.imread loop, including the clip's
    pulled-back last window (v2ce.py:153-154)."""
    frames, paths = _write_pngs(str(tmp_path), n_frames)
    starts, _mode = drv.window_schedule(n_frames, 16)
    read, close = drv._window_reader(paths, None, starts, 16)
    try:
        for k, st in enumerate(starts):
            got = read(k)
            want = drv._read_window(paths, None, int(st), 16)
            assert got.dtype == np.uint8 and np.array_equal(got, want) and np.array_equal(got, frames[st:st + 17])
    finally:
        close()


def test_window_reader_keeps_the_serial_path_for_negative_starts(tmp_path):
    """A 16-frame clip schedules its only window at start -1 (the reference's slice then yields one frame, SURVEY.md
    F8a): such schedules are not prefetched, so the behaviour stays that of the reference's own loop."""
    frames, paths = _write_pngs(str(tmp_path), 16)
    starts, mode = drv.window_schedule(16, 16)
    assert list(starts) == [-1] and mode == 15
    read, close = drv._window_reader(paths, None, starts, 16)
    assert np.array_equal(read(0), drv._read_window(paths, None, -1, 16))
    close()


def test_frame_prefetcher_reports_unreadable_files(tmp_path):
    frames, paths = _write_pngs(str(tmp_path), 40)
    with open(paths[20], 'wb') as f:
        f.write(b'not a png')
    starts, _ = drv.window_schedule(40, 16)
    read, close = drv._window_reader(paths, None, starts, 16)
    try:
        read(0)
        with pytest.raises(FileNotFoundError):
            read(1)
    finally:
        close()


def test_batches_generator_uses_prefetch_and_matches_preprocessing(tmp_path):
    frames, paths = _write_pngs(str(tmp_path), 40, h=26, w=40)
    got = list(drv._batches(paths, None, 16, 26, 2))
    starts, _ = drv.window_schedule(40, 16)
    want = [drv.image_pre_processing(frames[s:s + 17], 26) for s in starts]
    flat = [u for units, _last in got for u in units]
    assert [last for _u, last in got] == [False, True] and len(flat) == len(want)
    for a, b in zip(flat, want):
        assert a.shape == b.shape and bool((a == b).all())


def test_u8_window_batches_center_crop(tmp_path):
    """The raw-uint8 batching of stream_clip (native-resolution clips): center crop of v2ce.py:78, reference batching."""
    frames, paths = _write_pngs(str(tmp_path), 49, h=26, w=40)
    got = list(drv._window_batches_u8(paths, None, 16, 2, None, 30))
    assert [tuple(x.shape) for x, _ in got] == [(2, 17, 26, 30), (1, 17, 26, 30)]
    assert [last for _, last in got] == [False, True]
    flat = np.concatenate([x.numpy() for x, _ in got], axis=0)
    for k, st in enumerate(drv.window_schedule(49, 16)[0]):
        assert np.array_equal(flat[k], frames[st:st + 17, :, 20 - 15:20 + 15])


def test_background_generator_order_errors_and_early_stop():
    import threading
    import time

    def gen(n, fail_at=None, log=None):
        try:
            for i in range(n):
                if i == fail_at:
                    raise ValueError(f'boom at {i}')
                time.sleep(0.001)
                yield i, threading.current_thread().name
        finally:
            if log is not None:
                log.append('closed')

    items = list(drv._background(gen(20)))
    assert [i for i, _ in items] == list(range(20))
    assert all(name == 'v2ce-batches' for _, name in items)        # produced off the consumer's thread
    got = []
    with pytest.raises(ValueError, match='boom at 3'):
        for i, _ in drv._background(gen(10, fail_at=3)):
            got.append(i)
    assert got == [0, 1, 2]
    log = []
    it = drv._background(gen(1000, log=log), depth=2)
    assert next(it)[0] == 0
    it.close()                                                     # consumer leaves early: the producer must stop
    deadline = time.time() + 5
    while not log and time.time() < deadline:
        time.sleep(0.01)
    assert log == ['closed']
    assert not [t for t in threading.enumerate() if t.name == 'v2ce-batches' and t.is_alive()]


def test_background_batches_equal_inline_batches(tmp_path):
    frames, paths = _write_pngs(str(tmp_path), 49, h=26, w=40)
    inline = list(drv._batches(paths, None, 16, 26, 2))
    threaded = list(drv._background(drv._batches(paths, None, 16, 26, 2)))
    assert len(inline) == len(threaded)
    for (a, la), (b, lb) in zip(inline, threaded):
        assert la == lb and bool((a == b).all())


def test_percentile_from_order_statistics_equals_numpy():
    """Host half of the exact percentile (v2ce.py:262-264): the device hands back the two neighbouring order statistics,
    the host interpolates like np.percentile -- also for the gray preview, whose array is every value three times."""
    from v2ce_toolbox_b200 import event_frames as ef
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 7, 100, 4097):
        v = np.sort((rng.random(n) * 12).astype(np.float32))
        for mult in (1, 3):
            full = np.sort(np.repeat(v, mult)).astype(np.float64)
            for q in (0, 1, 37, 50, 90, 98, 99, 100):
                vi = (n * mult - 1) * (q / 100.0)
                lo = int(np.floor(vi))
                hi = min(lo + 1, n * mult - 1)
                got = ef.percentile_from_order_statistics(n, mult, q, np.float32(full[lo]), np.float32(full[hi]))
                assert got == np.percentile(full, q), (n, mult, q)


@pytest.mark.parametrize('fps', [24, 25, 30, 50, 60, 120, 240, 1000])
def test_ldati_params_constants_equal_the_oracle(fps):
    """The scalar constants the product hands to the kernels (ldati.make_params) against the oracle's (oracle.Consts,
    pinned to the reference): every reciprocal, step and bin start, bit for bit, plus the sort-key window."""
    from oracle import ldati_oracle as lo
    from v2ce_toolbox_b200 import ldati
    p = ldati.make_params(3, 8, 12, fps=fps, t0=0, seed=5, frame_base=2, flavor='cpu', device='cpu')
    k = lo.Consts(fps, 0, 'cpu')
    f32 = np.float32
    assert (p.fps64, p.nbins64, p.r_fps64, p.r_nbins64) == (float(fps), 9.0, k.r_fps64, k.r_c64)
    for got, want in ((p.r_fps32, k.r_fps32), (p.r_nbins32, k.r_c32), (p.vs32, k.vs32), (p.inv_vs32, k.inv_vs32),
                      (p.vs2_32, f32(k.vs2)), (p.r_vs2_32, k.r_vs2_32), (p.r6_32, k.r6_32), (p.eps6, k.eps6),
                      (p.eps8, k.eps8), (p.six32, f32(6)), (p.fps32, f32(fps)), (p.nbins32, f32(9))):
        assert f32(got).view(np.uint32) == f32(want).view(np.uint32)
    got_bs = np.array(p.binstart_t0_32[:9], dtype=np.float32)
    assert np.array_equal(got_bs.view(np.uint32), k.binstart_t0_32.view(np.uint32))
    one_bin = int(np.ceil(1e6 / fps / 9))
    for c in range(9):
        assert p.bin_base_us[c] == int(np.floor(float(k.binstart_t0_32[c]) * 1e6))
        # every in-contract timestamp of bin c lands inside the key window [1, 2^key_bits)
        assert p.key_span >= one_bin + 2 * ldati.KEY_BIAS
    assert (p.true_div, p.multi_events, p.bidirectional, p.pooling, p.frame_base, p.seed) == (1, 1, 0, 0, 2, 5)
    wide = ldati.make_params(3, 8, 12, fps=fps, flavor='cpu', device='cpu', bidirectional=True,
                             additional_events_strategy='random')
    assert wide.key_span >= 1_000_000 and wide.bin_base_us[0] < 0 and wide.multi_events == 2 and wide.bidirectional == 1


def test_split_frames_and_status_check():
    from v2ce_toolbox_b200 import V2ceError, ldati
    ev = np.zeros(10, ldati.EVENT_DTYPE)
    ev['timestamp'] = np.arange(10)
    seg = np.zeros((3, 9), np.int64)
    seg[0, 0], seg[0, 8], seg[2, 4] = 2, 3, 5
    parts = ldati.split_frames(ev.view(np.uint8), seg)
    assert [len(x) for x in parts] == [5, 0, 5] and parts[2]['timestamp'][0] == 5 and isinstance(parts[0], np.recarray)
    ldati.check_status(np.array([0, 3, 0, 0], np.int32))                 # NaN timestamps alone are legal (INT64_MIN)
    with pytest.raises(V2ceError):
        ldati.check_status(np.array([1, 0, 0, 0], np.int32))


def test_v2ce3d_checkpoint_handling_without_a_device():
    """load_state_dict accepts the reference's layout (BatchNorm's num_batches_tracked entries are dropped), nothing is
    built before a CUDA device is named, and there is no CPU path."""
    import torch
    from oracle import synth
    from v2ce_toolbox_b200 import V2ceError
    from v2ce_toolbox_b200.scripts.v2ce_3d import V2ce3d
    sd = synth.make_state_dict(0, 'reference')
    sd['UNet.encoders.0.bn1.num_batches_tracked'] = torch.tensor(7)
    m = V2ce3d()
    res = m.load_state_dict({k: v.double() for k, v in sd.items()})       # nn.Module's result type
    assert list(res.missing_keys) == [] and list(res.unexpected_keys) == []
    with pytest.raises(V2ceError):
        V2ce3d().load_state_dict({**sd, 'not_a_unet_key': torch.zeros(1)})
    assert V2ce3d().load_state_dict({**sd, 'not_a_unet_key': torch.zeros(1)}, strict=False).unexpected_keys == ['not_a_unet_key']
    got = m.state_dict()
    assert 'UNet.encoders.0.bn1.num_batches_tracked' not in got
    assert set(got) == {k for k in sd if not k.endswith('num_batches_tracked')}
    assert all(v.dtype == torch.float32 and v.device.type == 'cpu' for v in got.values())
    assert m.eval() is m and m.to(None) is m
    with pytest.raises(V2ceError):
        m.to('cpu')
    with pytest.raises(V2ceError):
        m(torch.zeros(1, 16, 2, 8, 8))                      # CPU tensor: no fallback
    with pytest.raises(V2ceError):
        V2ce3d(in_channels=3)
