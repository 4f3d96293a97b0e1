"""The product's host-side driver (v2ce_toolbox_b200/v2ce.py: pre-processing, window schedule, center crop, pano tiles,
batching, merge -- /root/reference/v2ce.py:45-239) against the UNMODIFIED reference driver, live, on CPU.

Both drivers are handed the same stand-in for the network (a cheap deterministic map (b,L,2,H,W) -> (b,L,20,H,W) that
also depends on the call index, like the real model's spectral-norm state) and the same in-memory clip, with
``Tensor.cuda`` shimmed to the identity; every float of the merged voxel grid must agree.  Runs only where
/root/reference is mounted."""
import numpy as np
import pytest
import torch

from oracle import ref_harness as rh, synth

pytestmark = pytest.mark.skipif(not rh.available(), reason='/root/reference is not mounted')


class StandIn(torch.nn.Module):
    """Deterministic, position- and call-dependent stand-in for V2ce3d."""

    def __init__(self):
        super().__init__()
        self.calls = 0

    def forward(self, x):
        self.calls += 1
        b, L, c, H, W = x.shape
        assert c == 2
        ramp = torch.arange(20, dtype=torch.float32).view(1, 1, 20, 1, 1) * 0.25
        pos = torch.arange(W, dtype=torch.float32).view(1, 1, 1, 1, W) * 1e-3 + \
            torch.arange(L, dtype=torch.float32).view(1, L, 1, 1, 1) * 1e-2
        return x.repeat_interleave(10, dim=2) * (1.0 + 0.125 * self.calls) + ramp + pos


CASES = [('center', 17, 1, 26, 40, 36), ('center', 20, 2, 26, 40, 36), ('center', 33, 1, 26, 40, 36),
         ('center', 36, 4, 26, 40, 36), ('center', 49, 3, 26, 40, 40), ('center', 18, 2, 26, 40, 30),
         ('center', 40, 2, 52, 80, 36),                       # resized 52x80 -> 26x40, then cropped
         ('pano', 18, 1, 26, 52, 20), ('pano', 20, 2, 26, 40, 20), ('pano', 35, 2, 26, 47, 20),
         ('pano', 33, 1, 52, 104, 20)]                        # resized, 3 tiles, the last one pulled back


@pytest.mark.parametrize('infer_type,n_frames,batch_size,H,W,width', CASES)
def test_video_to_voxels_equals_reference_driver(infer_type, n_frames, batch_size, H, W, width):
    from v2ce_toolbox_b200 import v2ce as drv
    frames = synth.make_video(n_frames, H, W, seed=n_frames)
    height = 26
    with rh.cpu_cuda_shims():
        want = rh.main_module().video_to_voxels(StandIn(), vidcap=rh.FakeVideoReader(frames), infer_type=infer_type,
                                                seq_len=16, width=width, height=height, batch_size=batch_size)
        got = drv.video_to_voxels(StandIn(), vidcap=rh.FakeVideoReader(frames), infer_type=infer_type, seq_len=16,
                                  width=width, height=height, batch_size=batch_size)
    want, got = np.asarray(want), np.asarray(got)
    assert got.shape == want.shape and got.shape[0] == n_frames - 1
    assert np.array_equal(got.astype(np.float32).view(np.uint32), want.astype(np.float32).view(np.uint32))


def test_window_schedule_and_tiles_equal_reference_arithmetic():
    """The closed forms the product uses against the reference's inline arithmetic (v2ce.py:149-154, 103-111)."""
    from v2ce_toolbox_b200 import v2ce as drv
    for n in range(2, 120):
        starts, mode = drv.window_schedule(n, 16)
        sequence_num = int(np.ceil((n - 1) / 16))
        ref = [i * 16 for i in range(sequence_num)]
        if (n - 1) % 16 != 0:
            ref[-1] = ref[-1] - (16 - (n - 1) % 16)
        assert list(starts) == ref and mode == (n - 1) % 16
    for total, width in ((462, 346), (1920, 346), (692, 346), (346, 346), (52, 20), (40, 20), (47, 20)):
        tiles = drv.pano_tiles(total, width)
        assert tiles[0][0] == 0 and tiles[-1][1] == total
        # a remainder of 0 with total % 346 != 0 keeps the whole pulled-back tile (`out[..., -0:]` in the reference,
        # v2ce.py:120-127): the driver comparison above ('pano', 20, 2, 26, 40, 20) covers that quirk end to end
        assert sum((k if k else width) for _, _, k in tiles) == total
