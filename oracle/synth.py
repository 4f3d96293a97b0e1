"""Seeded synthetic inputs (state dicts, videos, voxels) live in the repo-root module ``synth_inputs`` so that
``bench.py`` and ``tools/`` can build their workloads without importing anything under ``oracle/``; tests and the
golden generator keep reaching them under this name."""
from synth_inputs import *            # noqa: F401,F403
from synth_inputs import layer_table, make_state_dict, make_video, make_voxels, order_like   # noqa: F401
