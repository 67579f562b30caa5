"""Host-side logic of the drop-in class that needs no GPU: window clamping, batching,
LR synthesis, image IO conventions, shard arithmetic (model/pfnl.py:203-332)."""
import os

import numpy as np
import pytest

from oracle import pfnl_ref as R


def _model():
    from pfnl_b200 import PFNL
    return PFNL()


def test_attributes_match_reference(built_lib):
    m = _model()
    # model/pfnl.py:21-37
    assert (m.num_frames, m.scale, m.in_size, m.gt_size) == (7, 4, 32, 128)
    assert m.eval_in_size == [128, 240] and m.batch_size == 16 and m.eval_basz == 4
    assert m.save_dir == './checkpoint/pfnl' and m.log_dir == './pfnl.txt'
    for name in ("forward", "test_video_truth", "test_video_lr", "testvideos", "testvideo", "eval_mse"):
        assert callable(getattr(m, name))


def test_window_list_matches_reference_clamping(built_lib):
    m = _model()
    lrs = np.arange(5, dtype=np.float32).reshape(5, 1, 1, 1) * np.ones((5, 2, 2, 3), np.float32)
    wl = m._window_list(lrs)
    assert wl.shape == (5, 7, 2, 2, 3)
    for i in range(5):
        assert wl[i, :, 0, 0, 0].astype(int).tolist() == R.window_indices(5, i)


def test_downsample_4d_matches_independent_restatement(built_lib):
    """DownSample_4D (utils.py:169-192): REFLECT pad 6 + 13x13 Gaussian + stride 4, per channel."""
    from pfnl_b200.model import downsample_4d, gkern
    from scipy.ndimage import correlate
    rng = np.random.default_rng(0)
    x = rng.random((2, 32, 40, 3)).astype(np.float32)
    y = downsample_4d(x, 4)
    assert y.shape == (2, 8, 10, 3)
    k = gkern(13, 1.6)
    assert abs(k.sum() - 1.0) < 1e-3 and k.shape == (13, 13)
    ref = np.stack([np.stack([correlate(x[n, :, :, c].astype(np.float64), k, mode="mirror")[::4, ::4]
                              for c in range(3)], -1) for n in range(2)])
    np.testing.assert_allclose(y, ref, atol=1e-5)


def test_imsave_imread_roundtrip_rgb(tmp_path, built_lib):
    from pfnl_b200.model import cv2_imread, cv2_imsave
    import cv2
    img = np.zeros((4, 5, 3), np.uint8)
    img[..., 0] = 200  # red in RGB
    p = str(tmp_path / "a.png")
    cv2_imsave(p, img)
    np.testing.assert_array_equal(cv2_imread(p), img)
    assert cv2.imread(p)[0, 0].tolist() == [0, 0, 200]  # stored BGR (utils.py:362-366)


def test_shard_ranges_cover_everything(built_lib):
    from pfnl_b200.dist import shard_range
    for n in (1, 7, 16, 128, 129):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
