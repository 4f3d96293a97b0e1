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
