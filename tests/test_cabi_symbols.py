"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol
include/v2ce_b200.h declares; compute calls fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'v2ce_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(v2ce_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from v2ce_toolbox_b200 import _lib, build
    build.build_library()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/v2ce_b200.h but not exported'
    assert set(_lib.EXPORTED_SYMBOLS) == set(names)
    assert _lib.load().v2ce_version() >= 100


def test_no_cpu_fallback():
    from v2ce_toolbox_b200 import V2ceError
    from v2ce_toolbox_b200.scripts.LDATI import sample_voxel_statistical
    from v2ce_toolbox_b200.scripts.v2ce_3d import V2ce3d
    with pytest.raises(V2ceError):
        sample_voxel_statistical(torch.zeros(1, 2, 10, 4, 4))
    with pytest.raises(V2ceError):
        V2ce3d().to('cpu')
    if not torch.cuda.is_available():
        from v2ce_toolbox_b200 import _lib
        lib = _lib.load()
        assert lib.v2ce_device_check(0, None, None, None) != 0
        assert lib.v2ce_last_error()


def test_cli_flags_match_reference():
    from v2ce_toolbox_b200.v2ce import build_parser
    p = build_parser()
    d = vars(p.parse_args(['-i', 'x.mp4']))
    assert (d['fps'], d['seq_len'], d['ceil'], d['upper_bound_percentile'], d['out_folder'], d['infer_type'],
            d['model_path'], d['max_frame_num'], d['width'], d['height'], d['batch_size'], d['stage2_batch_size']) == \
           (30, 16, 10, 98, './output', 'center', './weights/v2ce_3d.pt', 1800, 346, 260, 1, 24)
    assert d['write_event_frame_video'] is True and d['vis_keep_polarity'] is True
    d = vars(p.parse_args(['-f', 'dir', '--write_event_frame_video', 'false', '-t', 'pano', '-b', '4', '-u', '95']))
    assert d['write_event_frame_video'] is False and d['infer_type'] == 'pano' and d['batch_size'] == 4


def test_host_schedule_matches_oracle():
    from oracle import pipeline_oracle as po
    from v2ce_toolbox_b200 import v2ce as drv
    import numpy as np
    for n in (17, 18, 33, 321, 600):
        a, ma = drv.window_schedule(n)
        b, mb = po.window_starts(n)
        assert ma == mb and np.array_equal(a, b)
    for w in (346, 462, 692, 1920):
        assert drv.pano_tiles(w) == po.pano_tiles(w)
    from oracle import synth
    fr = synth.make_video(5, 40, 52, seed=1)
    assert np.array_equal(drv.image_pre_processing(fr, 40).numpy(), po.preprocess(fr, 40))
    assert drv.frame_offset_us(7, 30) == po.frame_offset_us(7, 30)


def test_ldati_params_mirror_matches_the_c_struct():
    """The ctypes mirror of v2ce_ldati_params must have the C struct's size, and its LAST fields must land where the
    library reads them: the library's own argument check accepts a valid struct and names the field we then corrupt."""
    import ctypes
    from v2ce_toolbox_b200 import _lib, ldati
    lib = _lib.load()
    assert lib.v2ce_ldati_params_size() == ctypes.sizeof(_lib.LdatiParams)
    p = ldati.make_params(2, 8, 12, fps=30, flavor='cpu', device='cpu', additional_events_strategy='slope',
                          bidirectional=True, pooling_type='avg', pooling_kernel_size=5)
    assert lib.v2ce_ldati_params_validate(ctypes.byref(p)) == 0
    assert (p.multi_events, p.bidirectional, p.pooling, p.pooling_kernel_size) == (1, 1, 2, 5)
    p.pooling_kernel_size = 4
    assert lib.v2ce_ldati_params_validate(ctypes.byref(p)) != 0 and b'pooling_kernel_size' in lib.v2ce_last_error()
    p.pooling_kernel_size, p.pooling = 5, 7
    assert lib.v2ce_ldati_params_validate(ctypes.byref(p)) != 0 and b'pooling must be' in lib.v2ce_last_error()
    p.pooling, p.bidirectional = 0, 3
    assert lib.v2ce_ldati_params_validate(ctypes.byref(p)) != 0 and b'bidirectional' in lib.v2ce_last_error()
    p.bidirectional, p.multi_events = 0, 9
    assert lib.v2ce_ldati_params_validate(ctypes.byref(p)) != 0 and b'multi_events' in lib.v2ce_last_error()
    p.multi_events, p.key_span = 1, 0
    assert lib.v2ce_ldati_params_validate(ctypes.byref(p)) != 0 and b'key_span' in lib.v2ce_last_error()


def test_image_units_argument_checks_run_without_a_gpu():
    """The host-side checks of v2ce_image_units fire before any launch: binding arity and error text, on CPU."""
    from v2ce_toolbox_b200 import _lib
    lib = _lib.load()
    assert lib.v2ce_image_units(None, 1, 17, 260, 346, 260, 346, None, None) != 0
    assert b'NULL' in lib.v2ce_last_error()
    import ctypes
    one, out = (ctypes.c_uint8 * 16)(), (ctypes.c_float * 16)()
    assert lib.v2ce_image_units(one, 1, 1, 4, 4, 4, 4, out, None) != 0 and b'window geometry' in lib.v2ce_last_error()
    assert lib.v2ce_image_units(one, 1, 2, 1, 4, 4, 4, out, None) != 0 and b'frame geometry' in lib.v2ce_last_error()


def test_ldati_workspace_queries_are_host_only_and_consistent():
    """Workspace sizing runs without a device: the count workspace grows by exactly the per-pixel-bin count planes when
    the slope is fitted on pooled counts, the emit workspace grows with the event count and with 64-bit elements."""
    import ctypes
    from v2ce_toolbox_b200 import _lib, ldati
    lib = _lib.load()

    def count_bytes(**kw):
        p = ldati.make_params(24, 260, 346, flavor='cpu', device='cpu', **kw)
        n = ctypes.c_size_t()
        assert lib.v2ce_ldati_count_workspace_bytes(ctypes.byref(p), ctypes.byref(n)) == 0
        return n.value

    def emit_bytes(total, **kw):
        p = ldati.make_params(24, 260, 346, flavor='cpu', device='cpu', **kw)
        n = ctypes.c_size_t()
        assert lib.v2ce_ldati_emit_workspace_bytes(ctypes.byref(p), total, ctypes.byref(n)) == 0
        return n.value

    base = count_bytes()
    planes = 24 * 2 * 9 * 260 * 346 * 4
    assert 0 < base < 64 << 20
    assert planes <= count_bytes(pooling_type='weighted') - base < planes + 4096
    assert count_bytes(pooling_type='avg', additional_events_strategy='none') == base      # pooling only feeds 'slope'
    small, big = emit_bytes(1_000_000), emit_bytes(100_000_000)
    assert 8 * 1_000_000 <= small < big and big >= 8 * 100_000_000
    assert emit_bytes(1_000_000, additional_events_strategy='random') >= small + 8 * 1_000_000 - 4096   # 64-bit elements
    p = ldati.make_params(24, 260, 346, flavor='cpu', device='cpu')
    n = ctypes.c_size_t()
    assert lib.v2ce_ldati_emit_workspace_bytes(ctypes.byref(p), 1 << 31, ctypes.byref(n)) != 0
    assert b'split the frames' in lib.v2ce_last_error()


def test_entry_points_fail_loudly_without_a_device():
    """No CUDA device (this test is skipped where one exists): the library reports an error code and a message; it never
    crashes and there is nothing to fall back to."""
    import ctypes
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    from v2ce_toolbox_b200 import _lib
    lib = _lib.load()
    a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.v2ce_device_check(0, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)) < 0 and lib.v2ce_last_error()
    h = ctypes.c_void_p()
    assert lib.v2ce_model_create(ctypes.byref(h), 0) < 0 and not h.value
    assert lib.v2ce_ef_accumulate(None, 1, 4, 4, 1, None, None) < 0 and b'NULL' in lib.v2ce_last_error()
    with pytest.raises(_lib.V2ceError):
        _lib.check(lib.v2ce_model_create(ctypes.byref(h), 0))


def test_baseline_params_mirror_matches_the_c_struct():
    import ctypes
    from v2ce_toolbox_b200 import _lib
    from v2ce_toolbox_b200.sample_methods import random_even_sample as rs
    lib = _lib.load()
    assert lib.v2ce_baseline_params_size() == ctypes.sizeof(_lib.BaselineParams)
    p = rs.make_params(2, 8, 12, fps=30, mode='even', flavor='cpu', device='cpu')
    n = ctypes.c_size_t()
    assert lib.v2ce_baseline_count_workspace_bytes(ctypes.byref(p), ctypes.byref(n)) == 0 and n.value > 0
    assert lib.v2ce_baseline_emit_workspace_bytes(ctypes.byref(p), 1000, ctypes.byref(n)) == 0 and n.value >= 16000
    p.mode = 5                                            # the LAST fields land where the library reads them
    assert lib.v2ce_baseline_count_workspace_bytes(ctypes.byref(p), ctypes.byref(n)) != 0 and b'mode' in lib.v2ce_last_error()
    p.mode, p.key_span = 1, 0
    assert lib.v2ce_baseline_count_workspace_bytes(ctypes.byref(p), ctypes.byref(n)) != 0 and b'key_span' in lib.v2ce_last_error()
    assert lib.v2ce_ts_diff_workspace_bytes(346, 260, 1000, ctypes.byref(n)) == 0 and n.value > 346 * 260 * 2 * 12
