"""Second, independent implementations of the TF-1.12 op semantics the oracle restates, taken
from libraries in this image that share no code with oracle/pfnl_ref.py.  The reference itself holds no vectors and
TF 1.12 cannot run here, so the oracle stays "parity unpinned" by the reference; these checks bound the
risk that the restatement and the CUDA kernels share a misreading of an op.

  op (reference)                               independent implementation
  legacy ResizeBicubic x4 (pfnl.py:63)         torch grid_sample(mode='bicubic', border): Keys cubic A = -0.75 at
                                               src = dst / 4 (no half-pixel shift), tap indices clamped
  depth_to_space / space_to_depth (DCR)        einops.rearrange patterns; torch pixel_shuffle must differ (CRD)
  Conv2D 'same', cross-correlation (A.1)       scipy.signal.correlate2d(mode='same', zero fill) per channel pair
  leaky_relu(alpha = 0.2)                      torch.nn.functional.leaky_relu
  NonLocalBlock nltype=1 (utils.py:18-71)      torch softmax attention with the two 1x1 convs as matmuls
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pfnl_ref as R


def test_bicubic_vs_torch_grid_sample():
    rng = np.random.default_rng(0)
    img = rng.random((2, 9, 11, 3))
    ours = R.resize_bicubic(img, 36, 44)
    n, h, w, c = img.shape
    # sample positions of the legacy kernel: src = dst * (in/out) = dst / 4, align_corners=False off => no 0.5 shift
    ys = torch.arange(4 * h, dtype=torch.float64) / 4.0
    xs = torch.arange(4 * w, dtype=torch.float64) / 4.0
    # grid_sample with align_corners=True maps [-1, 1] to pixel centres 0 .. size-1
    gy = 2.0 * ys / (h - 1) - 1.0
    gx = 2.0 * xs / (w - 1) - 1.0
    grid = torch.stack(torch.meshgrid(gy, gx, indexing="ij")[::-1], -1)[None].expand(n, -1, -1, -1)
    t = torch.from_numpy(img).permute(0, 3, 1, 2)
    ref = F.grid_sample(t, grid, mode="bicubic", padding_mode="border", align_corners=True).permute(0, 2, 3, 1).numpy()
    # interior: identical kernels.  (At the far border the legacy kernel clamps tap INDICES while grid_sample clamps
    # the sample position first - rows/columns whose taps leave the image are compared separately below.)
    inner = np.abs(ours - ref)[:, : 4 * (h - 2), : 4 * (w - 2)]
    assert inner.max() < 1e-12
    # every 4th output pixel is an input pixel, everywhere (both conventions agree on that)
    np.testing.assert_array_equal(ours[:, ::4, ::4], img)


def test_bicubic_border_taps_by_direct_formula():
    """Direct evaluation of the Keys kernel (A = -0.75) with clamped tap indices, written from the formula only."""
    def keys(x, a=-0.75):
        x = abs(x)
        if x <= 1:
            return (a + 2) * x ** 3 - (a + 3) * x ** 2 + 1
        if x < 2:
            return a * x ** 3 - 5 * a * x ** 2 + 8 * a * x - 4 * a
        return 0.0
    rng = np.random.default_rng(1)
    row = rng.random(7)
    img = np.broadcast_to(row[None, None, :, None], (1, 5, 7, 1)).copy()
    ours = R.resize_bicubic(img, 20, 28)[0, 0, :, 0]
    for o in range(28):
        k, d = divmod(o, 4)
        d /= 4.0
        ref = sum(keys(d - j) * row[min(max(k + j, 0), 6)] for j in (-1, 0, 1, 2))
        assert abs(ours[o] - ref) < 1e-6, (o, ours[o], ref)   # the legacy kernel reads a 1024-entry table


@pytest.mark.parametrize("b,co", [(2, 21), (2, 12), (2, 3), (4, 3), (2, 1)])
def test_dcr_reorders_vs_einops(b, co):
    einops = pytest.importorskip("einops")
    rng = np.random.default_rng(b * 10 + co)
    x = rng.random((2, 3, 5, b * b * co)).astype(np.float32)
    d2s = einops.rearrange(x, "n h w (b1 b2 c) -> n (h b1) (w b2) c", b1=b, b2=b)
    np.testing.assert_array_equal(R.depth_to_space(x, b), d2s)
    np.testing.assert_array_equal(R.periodic_shuffle(x, b, co), d2s)       # modules/ps.py:_PS
    y = rng.random((2, 3 * b, 5 * b, co)).astype(np.float32)
    s2d = einops.rearrange(y, "n (h b1) (w b2) c -> n h w (b1 b2 c)", b1=b, b2=b)
    np.testing.assert_array_equal(R.space_to_depth(y, b), s2d)
    if co > 1:   # PyTorch's pixel_shuffle is CRD: a different permutation
        ps = F.pixel_shuffle(torch.from_numpy(x).permute(0, 3, 1, 2), b).permute(0, 2, 3, 1).numpy()
        assert not np.array_equal(ps, d2s)


@pytest.mark.parametrize("k", [1, 3, 5])
def test_conv_same_vs_scipy_correlate(k):
    sig = pytest.importorskip("scipy.signal")
    rng = np.random.default_rng(k)
    x = rng.standard_normal((1, 7, 9, 3))
    w = rng.standard_normal((k, k, 3, 4))
    bias = rng.standard_normal(4)
    ours = R.conv2d_same(x, w, bias, act=False)
    ref = np.zeros((7, 9, 4))
    for o in range(4):
        for c in range(3):
            ref[..., o] += sig.correlate2d(x[0, ..., c], w[..., c, o], mode="same", boundary="fill", fillvalue=0.0)
        ref[..., o] += bias[o]
    assert np.abs(ours[0] - ref).max() < 1e-12
    act = R.conv2d_same(x, w, bias, act=True)
    np.testing.assert_allclose(act[0], F.leaky_relu(torch.from_numpy(ref), 0.2).numpy(), atol=1e-12)


def test_nonlocal_vs_torch_softmax_attention():
    rng = np.random.default_rng(5)
    n, h2, w2, c = 2, 5, 6, 84
    x = rng.random((n, h2, w2, c))
    wg, bg = rng.standard_normal((1, 1, c, c)) * 0.1, rng.standard_normal(c) * 0.1
    ww, bw = rng.standard_normal((1, 1, c, c)) * 0.1, rng.standard_normal(c) * 0.1
    ours = R.nonlocal_block(x, wg, bg, ww, bw)
    t = torch.from_numpy(x).reshape(n, h2 * w2, c)
    g = t @ torch.from_numpy(wg[0, 0]) + torch.from_numpy(bg)
    p = torch.softmax(t @ t.transpose(1, 2), dim=-1)       # embedded-Gaussian-free 'gaussian' mode: theta = phi = x
    ref = ((p @ g) @ torch.from_numpy(ww[0, 0]) + torch.from_numpy(bw)).reshape(n, h2, w2, c).numpy()
    assert np.abs(ours - ref).max() < 1e-10
