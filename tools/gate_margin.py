"""Regime-A (Xavier weights, outputs ~ +-100) max-abs error of the fp16x3 mode vs the fp64 oracle over several
input seeds, with the split-operand tcgen05 non-local block (default) - run again with PFNL_NL_FFMA=1 for
the fp32 CUDA-core non-local block."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
from oracle import pfnl_ref as R  # noqa: E402  (checker)
from pfnl_b200 import Engine  # noqa: E402

W = R.make_weights("A")
e = Engine(W, 0, "fp16x3", graphs=False)
errs = []
for seed in range(100, 108):
    x = R.make_input(1, 32, 32, seed=seed)
    ref = R.pfnl_forward(x, W, dtype=np.float64, backend="torch")
    y = e.forward(torch.from_numpy(x).cuda()).cpu().numpy()
    errs.append(float(np.abs(y - ref).max()))
print("NL", "ffma" if os.environ.get("PFNL_NL_FFMA") else "tcgen05-split", "regime A max-abs per seed:",
      " ".join(f"{v:.2e}" for v in errs), "| worst", f"{max(errs):.2e}")
