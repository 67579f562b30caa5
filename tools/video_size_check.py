"""Forward at real video sizes (SURVEY 7 'hard part 3': Vid4 'calendar' 720x576 HR -> 180x144 LR, L = 6480;
UDM10 1272x720 HR -> 318x180 LR, L = 14310): every precision runs, outputs agree, time per frame."""
import json
import statistics
import sys

import torch

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfnl_b200 import Engine, weights as WT  # noqa: E402

W = WT.xavier_init()
W = {k: (v * 0.1 if k.startswith("nlvsr/conv2_") and k.endswith("kernel") else v) for k, v in W.items()}  # trained-like
out = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for name, (h, w) in {"vid4_calendar_144x180": (144, 180), "udm10_180x318": (180, 318)}.items():
    x = torch.rand(1, 7, h, w, 3, device='cuda')
    ref = None
    row = {}
    for prec in ("fp32", "fp16x3", "fp16x3:phase", "fp16x3_nltc", "fp16"):
        e = Engine(W, 0, prec.split(":")[0], graphs=True)
        if prec.endswith(":phase"):
            e.set_flow(False)   # the two-launches-per-block kernels, for comparison with the dataflow kernel
        for _ in range(2):
            y = e.forward(x)
        ts = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            y = e.forward(x)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        if ref is None:
            ref = y.clone()
        row[prec] = {"ms_per_frame": statistics.median(ts), "max_abs_vs_fp32": float((y - ref).abs().max()),
                     "hr_px_per_s": 16 * h * w / (statistics.median(ts) / 1e3)}
        e.close()
        del e
    out[name] = row
print(json.dumps(out))
