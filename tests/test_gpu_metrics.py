"""CUDA luma metrics and the eval() loop against the CPU oracle, through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_ref as M
from oracle import pfnl_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(built_lib):
    from pfnl_b200 import Engine
    return Engine(R.make_weights("B"), device=0, precision="fp32", graphs=False)


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


@pytest.mark.parametrize("shape,border", [((5, 40, 56), 8), ((2, 17, 23), 0), ((1, 33, 19), 3), ((3, 128, 96), 8)])
@pytest.mark.parametrize("round_y", [False, True])
def test_msy_vs_oracle(engine, shape, border, round_y):
    f, h, w = shape
    rng = np.random.default_rng(h * 3 + w)
    a = rng.random((f, h, w, 3), dtype=np.float32)
    b = np.clip(a + rng.normal(0, 0.03, a.shape).astype(np.float32), -0.1, 1.1)    # exercises the clip
    got = engine.msy(cu(a), cu(b), 0.0, 1.0, border, round_y).cpu().numpy()
    ref = M.msy(a, b, 0, 1, border, round_y)
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)   # double sums in a different order


def test_avg_psnr_and_compute_psnr_wrappers(engine, built_lib):
    from pfnl_b200 import PFNL
    m = PFNL(weights=R.make_weights("B"), precision="fp32")
    rng = np.random.default_rng(3)
    t = rng.integers(0, 256, (9, 48, 64, 3)).astype(np.float32)
    p = np.clip(t + rng.normal(0, 4, t.shape), 0, 255).astype(np.float32)
    assert abs(m.avg_psnr(t, p) - M.avg_psnr(t, p)) < 1e-9                        # utils.py defaults: 0..255
    assert abs(m.avg_psnr(t / 255, p / 255, vmin=0, vmax=1, t_border=1, sp_border=4) -
               M.avg_psnr(t / 255, p / 255, 0, 1, 1, 4)) < 1e-9
    tu, pu = t.astype(np.uint8), np.round(p).astype(np.uint8)
    got = m.psnr_y(tu, pu)
    ref = [M.compute_psnr(tu[i], pu[i]) for i in range(9)]
    np.testing.assert_allclose(got, ref, rtol=1e-12)


@pytest.mark.parametrize("shape", [(3, 24, 31), (2, 11, 11), (1, 64, 75), (2, 130, 70)])
def test_ssim_vs_oracle(engine, shape):
    f, h, w = shape
    rng = np.random.default_rng(h + w)
    a = rng.integers(0, 256, (f, h, w, 3)).astype(np.uint8)
    noise = rng.normal(0, 12, a.shape)
    b = np.clip(a + noise, 0, 255).astype(np.uint8)
    got = engine.ssim_y(cu(a), cu(b), 0.0, 255.0).cpu().numpy()
    ref = np.array([M.ssim(a[i], b[i]) for i in range(f)])
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-12)
    same = engine.ssim_y(cu(a), cu(a), 0.0, 255.0).cpu().numpy()
    np.testing.assert_allclose(same, 1.0, atol=1e-12)


def test_metric_errors(engine):
    from pfnl_b200._lib import PfnlError
    a = cu(np.zeros((1, 10, 40, 3), np.float32))
    with pytest.raises(PfnlError):
        engine.ssim_y(a, a)                      # smaller than the 11x11 window
    with pytest.raises(PfnlError):
        engine.msy(a, a, 0.0, 1.0, 5, False)     # border eats the whole frame
    with pytest.raises(PfnlError):
        engine.msy(a, a, 1.0, 1.0, 0, False)     # vmax <= vmin


def test_eval_loop_matches_oracle_pipeline(tmp_path, built_lib):
    """PFNL.eval() (model/pfnl.py:94-149) on synthetic PNG sequences at a reduced eval_in_size: same clips,
    same LR synthesis, same MSE/PSNR and log line as the oracle pipeline run on the host."""
    import cv2
    from pfnl_b200 import PFNL
    from pfnl_b200.model import downsample_4d
    rng = np.random.default_rng(77)
    in_h, in_w = 16, 24
    fh, fw = in_h * 4 + 16 + 3, in_w * 4 + 16 + 5           # larger than crop + border, like real frames
    seqs = []
    for s, nframes in enumerate((50, 48, 20)):              # centres 15,47 | 15,47 | 15 -> 5 clips, 4 used
        d = tmp_path / f"seq{s}" / "truth"
        d.mkdir(parents=True)
        base = rng.integers(0, 256, (fh, fw, 3)).astype(np.float32)
        for i in range(nframes):
            img = np.clip(base + 20 * np.sin(i / 3.0) + rng.normal(0, 2, base.shape), 0, 255).astype(np.uint8)
            cv2.imwrite(str(d / f"{i:04d}.png"), img[:, :, ::-1])
        seqs.append(str(tmp_path / f"seq{s}"))
    lst = tmp_path / "filelist_val.txt"
    lst.write_text("\n".join(seqs) + "\n")
    W = R.make_weights("B")
    m = PFNL(weights=W, precision="fp32")
    m.eval_in_size = [in_h, in_w]
    m.eval_dir = str(lst)
    m.log_dir = str(tmp_path / "pfnl.txt")
    m.global_step = 1234
    psnr_avg, mse_avg = m.eval()
    # oracle pipeline on the host
    clips = []
    for s, nframes in enumerate((50, 48, 20)):
        files = sorted(os.listdir(tmp_path / f"seq{s}" / "truth"))
        for idx0 in range(15, nframes, 32):
            idx = np.clip(np.arange(idx0 - 3, idx0 + 4), 0, nframes - 1)
            gt = [cv2.imread(str(tmp_path / f"seq{s}" / "truth" / files[i]))[:, :, ::-1] for i in idx]
            clips.append(np.stack([g[8:in_h * 4 + 8, 8:in_w * 4 + 8].astype(np.float32) / 255.0 for g in gt]))
    assert len(clips) == 5
    gt = np.stack(clips[:4])
    lr = downsample_4d(gt.reshape(-1, in_h * 4, in_w * 4, 3), 4).reshape(4, 7, in_h, in_w, 3)
    sr = R.pfnl_forward(lr, W, dtype=np.float64)
    mse = R.mse_per_clip(sr, gt[:, 3:4])
    np.testing.assert_allclose(mse_avg, np.mean(mse, axis=0), rtol=2e-3)
    np.testing.assert_allclose(psnr_avg, np.mean(10 * np.log10(1.0 / mse), axis=0), atol=5e-3)
    line = open(m.log_dir).read().strip()
    assert line.startswith('{"Iter": 1234 , "PSNR": [') and line.endswith('}')
