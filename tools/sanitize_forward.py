"""Tiny forwards in every precision (+ the stage entry points) for compute-sanitizer runs."""
import sys

import torch

sys.path.insert(0, '/root/repo')
from pfnl_b200 import Engine, weights as WT  # noqa: E402

W = WT.xavier_init()
for prec in ("fp32", "fp16x3", "fp16x3_nltc", "fp16"):
    e = Engine(W, 0, prec, graphs=False)
    for shape in ((2, 16, 16), (1, 10, 14)):
        x = torch.rand(shape[0], 7, shape[1], shape[2], 3, device='cuda')
        y = e.forward(x)
        hr = torch.rand_like(y)
        m = e.mse(y, hr)
        q = e.quantize_u8(y)
        torch.cuda.synchronize()
    a = torch.rand(3, 24, 31, 3, device='cuda')
    e.msy(a, a * 0.9, 0.0, 1.0, 2, True)
    e.ssim_y(a, a * 0.9, 0.0, 1.0)
    e.gather_windows(torch.rand(5, 8, 8, 3, device='cuda'), 0, 5)
    e.downsample4(torch.rand(2, 33, 47, 3, device='cuda'))
    torch.cuda.synchronize()
    print(prec, 'ok', float(y.abs().max()), flush=True)
    e.close()
