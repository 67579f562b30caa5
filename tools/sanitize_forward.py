"""Small forwards in every precision, the stage entry points, the key-split non-local block, the pipelined host
feed and the metrics kernels - the workload of the compute-sanitizer runs (tools/gpurun/sanitize.sh).
    python tools/sanitize_forward.py [precision ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfnl_b200 import Engine, weights as WT  # noqa: E402

W = WT.xavier_init()
precs = sys.argv[1:] or ["fp32", "fp16x3", "fp16x3_nltc", "fp16"]
for prec in precs:
    e = Engine(W, 0, prec, graphs=False)
    for shape in ((2, 16, 16), (1, 10, 14)):
        x = torch.rand(shape[0], 7, shape[1], shape[2], 3, device='cuda')
        y = e.forward(x)
        hr = torch.rand_like(y)
        m = e.mse(y, hr)
        q = e.quantize_u8(y)
        torch.cuda.synchronize()
    if prec != "fp32":
        n, h, w = 1, 18, 12
        fr = torch.randn(n * 7, h, w, 64, device='cuda')
        e.pfrb(2, fr, n, h, w)
        e.convmerge1(fr, n, h, w)
        e.conv0(torch.randn(n, h, w, 21, device='cuda'))
        e.nonlocal_block(torch.rand(1, 600, 84, device='cuda'))      # 5 key tiles: key split + merge
        t, o = e.forward_host_submit(torch.rand(1, 7, 8, 8, 3))
        t2, o2 = e.forward_host_submit(torch.rand(1, 7, 8, 8, 3).double())
        e.forward_host_wait(t)
        e.forward_host_wait(t2)
        torch.cuda.synchronize()
    a = torch.rand(3, 24, 31, 3, device='cuda')
    e.msy(a, a * 0.9, 0.0, 1.0, 2, True)
    e.ssim_y(a, a * 0.9, 0.0, 1.0)
    e.gather_windows(torch.rand(5, 8, 8, 3, device='cuda'), 0, 5)
    e.downsample4(torch.rand(2, 33, 47, 3, device='cuda'))
    torch.cuda.synchronize()
    print(prec, 'ok', float(y.abs().max()), flush=True)
    e.close()
