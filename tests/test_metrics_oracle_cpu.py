"""Known-answer tests of the metric oracle (oracle/metrics_ref.py: utils.py:194-246, matlab/SSIM.m,
matlab/compute_psnr.m) - CPU only."""
import numpy as np

from oracle import metrics_ref as M


def test_luma_of_black_white_and_primaries():
    px = np.array([[[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255]]], np.float64)
    y = M.rgb2ycbcr(px, 255)[0, :, 0]
    np.testing.assert_allclose(y, [16.0, 235.0, 16 + 65.481, 16 + 128.553, 16 + 24.966], atol=1e-9)
    assert np.array_equal(M.matlab_y_uint8(px.astype(np.uint8))[0], [16, 235, 81, 145, 41])


def test_to_uint8_rounds_half_to_even_and_clips():
    x = np.array([-1.0, 0.0, 0.5 / 255, 1.5 / 255, 2.5 / 255, 1.0, 2.0], np.float32)
    got = M.to_uint8(x, 0, 1)
    # fp32 (x-vmin)/(vmax-vmin)*255 of k+0.5 ties: 0.5 -> 0, 1.5 -> 2, 2.5 -> 2 (numpy round half to even)
    assert got.tolist() == [0, 0, 0, 2, 2, 255, 255]


def test_gaussian_window_matches_fspecial():
    w = M.gaussian_window()
    assert w.shape == (11, 11) and abs(w.sum() - 1) < 1e-15
    assert np.allclose(w, w.T) and np.allclose(w, w[::-1, ::-1])
    # fspecial('gaussian', 11, 1.5) is separable: centre tap = 1 / (sum_x exp(-x^2 / 4.5))^2 = 0.0707622...
    s1 = sum(np.exp(-x * x / 4.5) for x in range(-5, 6))
    assert abs(w[5, 5] - 1.0 / s1 ** 2) < 1e-15 and abs(w[5, 5] - 0.0707622) < 1e-7
    assert abs(w[0, 0] - np.exp(-50 / 4.5) / s1 ** 2) < 1e-18


def test_ssim_closed_forms():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (24, 31, 3), dtype=np.uint8)
    assert abs(M.ssim(a, a) - 1.0) < 1e-12
    b = rng.integers(0, 256, (24, 31, 3), dtype=np.uint8)
    assert abs(M.ssim(a, b) - M.ssim(b, a)) < 1e-12 and M.ssim(a, b) < 0.2
    # two constant images: sigma terms vanish, SSIM = (2 y1 y2 + C1) / (y1^2 + y2^2 + C1)
    c1 = np.full((16, 16, 3), 100, np.uint8)
    c2 = np.full((16, 16, 3), 140, np.uint8)
    y1, y2 = M.matlab_y_uint8(c1)[0, 0], M.matlab_y_uint8(c2)[0, 0]
    C1 = (0.01 * 255) ** 2
    assert abs(M.ssim(c1, c2) - (2 * y1 * y2 + C1) / (y1 * y1 + y2 * y2 + C1)) < 1e-12
    assert M.ssim(a[:10], a[:10]) == -np.inf        # smaller than the window: SSIM.m returns -Inf


def test_psnr_closed_forms():
    a = np.full((3 + 4, 40, 40, 3), 0.5, np.float32)
    b = a.copy()
    b[..., 1] += 10 / 255.0                          # +10 grey levels on G -> luma offset 10 * 0.504129...
    want = 20 * np.log10(255.0 / (10 * 0.504129411764706))
    assert abs(M.avg_psnr(a, b, vmin=0, vmax=1) - want) < 1e-9
    m = M.msy(a, b, 0, 1, 8)
    assert m.shape == (7,) and np.allclose(m, (10 * 0.504129411764706) ** 2)
    u1 = np.full((20, 20, 3), 100, np.uint8)
    u2 = np.full((20, 20, 3), 110, np.uint8)
    d = M.matlab_y_uint8(u1)[0, 0] - M.matlab_y_uint8(u2)[0, 0]
    assert abs(M.compute_psnr(u1, u2) - 20 * np.log10(255.0 / abs(d))) < 1e-12
    # border crop: an error confined to the 8-pixel frame border is invisible to AVG_PSNR
    c = a.copy()
    c[:, :8] = 0
    with np.errstate(divide="ignore"):
        assert np.isinf(M.avg_psnr(a, c, vmin=0, vmax=1))
