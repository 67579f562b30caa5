"""Freezes oracle outputs as small fixtures (run from the repo root:
    python tests/golden/make_golden.py).
The reference has no golden vectors of its own and TensorFlow 1.12 cannot be
installed here, so these pin the *oracle* (oracle/pfnl_ref.py, numpy fp32 back-end) against
drift; the oracle itself is pinned by the known-answer tests in tests/test_oracle_kat.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pfnl_ref as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # whole forward, tiny frames (8x8 and a ragged 6x10), both weight regimes
    for regime in "AB":
        W = R.make_weights(regime)
        for (n, h, w) in [(1, 8, 8), (2, 6, 10)]:
            x = R.make_input(n, h, w, seed=1234)
            y = R.pfnl_forward(x, W, dtype=np.float32)
            y64 = R.pfnl_forward(x, W, dtype=np.float64)
            np.savez_compressed(os.path.join(HERE, f"forward_{regime}_{n}x{h}x{w}.npz"), x=x, y=y,
                                y64=y64.astype(np.float32))
    # the PR1 parity configuration: 1 clip x 7 x 32x32 (fp64-evaluated, stored fp32)
    for regime in "AB":
        W = R.make_weights(regime)
        x = R.make_input(1, 32, 32, seed=1234)
        y64 = R.pfnl_forward(x, W, dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, f"forward_{regime}_1x32x32.npz"), y64=y64.astype(np.float32))
    # non-local block alone, L = 60 (10x6 tokens, ragged) and 256
    W = R.make_weights("B")
    P = "nlvsr/nlblock_0/"
    for (hh, ww) in [(10, 6), (16, 16)]:
        rng = np.random.default_rng(77)
        t = rng.random((2, hh, ww, 84), dtype=np.float32)
        z = R.nonlocal_block(t.astype(np.float64), W[P + "g/g/kernel"].astype(np.float64), W[P + "g/g/bias"].astype(np.float64),
                             W[P + "w/w/kernel"].astype(np.float64), W[P + "w/w/bias"].astype(np.float64))
        np.savez_compressed(os.path.join(HERE, f"nonlocal_{hh}x{ww}.npz"), t=t, z=z.astype(np.float32))
    # bicubic x4 of a 5x7x3 image
    rng = np.random.default_rng(5)
    img = rng.random((1, 5, 7, 3), dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "bicubic_5x7.npz"), img=img, out=R.resize_bicubic(img, 20, 28))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
