import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def record_measurement(kind, **fields):
    """Append one measured figure (voxel deviations per shape, ...) to gpurun_out/measurements_r2.jsonl on the GPU box:
    the stated tolerances in the tests are set from these."""
    import json
    try:
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'measurements_r2.jsonl'), 'a') as f:
            f.write(json.dumps({'kind': kind, **fields}) + '\n')
    except OSError:
        pass


@pytest.fixture(scope='session')
def golden_meta():
    import json
    with open(os.path.join(GOLDEN, 'golden_meta.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + '_golden.npz'))
    return load
