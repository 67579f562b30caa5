"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement, in numpy float64, of the quality metrics the reference applies to the hot path's
output:

* `to_uint8`, `rgb2ycbcr`, `avg_psnr`  follow utils.py:194-246 (`_rgb2ycbcr`, `to_uint8`, `AVG_PSNR`);
* `matlab_y_uint8`, `compute_psnr`, `ssim` follow matlab/compute_psnr.m:1-18 and matlab/SSIM.m (the
  Wang et al. index on the luma channel with its default arguments: 11x11 Gaussian window of sigma
  1.5, 'valid' filtering, K = (0.01, 0.03), L = 255).  MATLAB's `rgb2ycbcr` on a uint8 image returns
  uint8 (round half away from zero, saturate) - restated from its documented behaviour; MATLAB
  itself is not available, so these two are PARITY UNPINNED and anchored by known-answer tests
  (identical images -> SSIM 1 / infinite PSNR, constant offsets -> closed forms).
"""
import numpy as np

_T = np.array([[0.256788235294118, 0.504129411764706, 0.097905882352941],
               [-0.148223529411765, -0.290992156862745, 0.439215686274510],
               [0.439215686274510, -0.367788235294118, -0.071427450980392]])
_O = np.array([16.0, 128.0, 128.0])


def to_uint8(x, vmin, vmax):
    """utils.py:211-214 (float32 arithmetic, np.round = half to even); returns float32 integers."""
    x = np.asarray(x).astype('float32')
    x = (x - np.float32(vmin)) / (np.float32(vmax) - np.float32(vmin)) * np.float32(255)
    return np.clip(np.round(x), 0, 255)


def rgb2ycbcr(img, max_val=255):
    """utils.py:194-209 for one [H,W,3] image."""
    o = _O / 255.0 if max_val == 1 else _O
    t = np.reshape(img, (img.shape[0] * img.shape[1], img.shape[2])).astype(np.float64)
    t = np.dot(t, _T.T) + o[None, :]
    return np.reshape(t, img.shape)


def luma(vid, vmin, vmax):
    """Y planes [F,H,W] exactly as AVG_PSNR builds them (utils.py:226-236)."""
    return np.stack([rgb2ycbcr(to_uint8(f, vmin, vmax), 255)[:, :, 0] for f in vid], 0)


def avg_psnr(vid_true, vid_pred, vmin=0, vmax=255, t_border=2, sp_border=8):
    """utils.py:216-246 with is_T_Y = is_P_Y = False."""
    d = luma(vid_true, vmin, vmax) - luma(vid_pred, vmin, vmax)
    n = d.shape[0]
    d = d[t_border:n - t_border, sp_border:d.shape[1] - sp_border, sp_border:d.shape[2] - sp_border]
    psnrs = [20 * np.log10(255. / np.sqrt(np.mean(np.power(d[t], 2)))) for t in range(d.shape[0])]
    return np.mean(np.asarray(psnrs))


def msy(vid_a, vid_b, vmin, vmax, sp_border, round_y=False):
    """Per-frame mean squared luma difference over the cropped frame (what pfnl_msy returns)."""
    ya, yb = luma(vid_a, vmin, vmax), luma(vid_b, vmin, vmax)
    if round_y:
        ya, yb = np.clip(np.floor(ya + 0.5), 0, 255), np.clip(np.floor(yb + 0.5), 0, 255)
    d = (ya - yb)[:, sp_border:ya.shape[1] - sp_border, sp_border:ya.shape[2] - sp_border]
    return np.mean(d * d, axis=(1, 2))


def matlab_y_uint8(img_u8):
    """Y channel of MATLAB rgb2ycbcr for a uint8 RGB image: uint8(round(T[0].rgb + 16))."""
    y = img_u8.astype(np.float64) @ _T[0] + 16.0
    return np.clip(np.floor(y + 0.5), 0, 255)


def compute_psnr(img1_u8, img2_u8):
    """matlab/compute_psnr.m:1-18 (boundarypixels = 0)."""
    d = matlab_y_uint8(img1_u8) - matlab_y_uint8(img2_u8)
    return 20 * np.log10(255.0 / np.sqrt(np.mean(d * d)))


def gaussian_window(size=11, sigma=1.5):
    r = np.arange(size) - (size - 1) / 2.0
    w = np.exp(-(r[:, None] ** 2 + r[None, :] ** 2) / (2.0 * sigma * sigma))
    return w / w.sum()


def _filter2_valid(w, img):
    win = np.lib.stride_tricks.sliding_window_view(img, w.shape)
    return np.einsum("hwij,ij->hw", win, w)


def ssim(img1_u8, img2_u8, K=(0.01, 0.03), L=255):
    """matlab/SSIM.m with two arguments: mean SSIM of the luma of two uint8 RGB images."""
    y1, y2 = matlab_y_uint8(img1_u8), matlab_y_uint8(img2_u8)
    if y1.shape[0] < 11 or y1.shape[1] < 11:
        return -np.inf
    w = gaussian_window()
    C1, C2 = (K[0] * L) ** 2, (K[1] * L) ** 2
    mu1, mu2 = _filter2_valid(w, y1), _filter2_valid(w, y2)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _filter2_valid(w, y1 * y1) - mu1_sq
    s2 = _filter2_valid(w, y2 * y2) - mu2_sq
    s12 = _filter2_valid(w, y1 * y2) - mu12
    m = ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return float(np.mean(m))
