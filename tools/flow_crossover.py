"""Dataflow kernel vs the two-launches-per-block phase kernels over (clips, height, width): where does each win?
The bench workload keeps the whole working set of a block (actA + actB + base) inside the 126 MB L2; a 180x318 frame
does not.  Prints one JSON object: {shape: {"units": U, "flow_ms": .., "phase_ms": ..}}.
Usage: python tools/flow_crossover.py [precision]   (PFNL_FLOW_L2_HINTS=0 to run without the L2 eviction hints)"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfnl_b200 import Engine, weights as WT  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
W = WT.xavier_init()
shapes = [(16, 32, 32), (32, 32, 32), (64, 32, 32), (128, 32, 32), (4, 64, 64), (8, 64, 64), (16, 64, 64), (1, 96, 96),
          (1, 128, 128), (1, 144, 180), (2, 144, 180), (1, 180, 318), (1, 270, 480)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
eng = {}
for mode in ("flow", "phase"):
    e = Engine(W, 0, prec, graphs=True)
    e.set_flow(mode == "flow")
    eng[mode] = e
out = {}
for (n, h, w) in shapes:
    x = torch.rand(n, 7, h, w, 3, device='cuda')
    row = {"units": n * ((w + 7) // 8) * ((h + 15) // 16)}
    ys = {}
    for mode, e in eng.items():
        for _ in range(2):
            y = e.forward(x)
        ts = []
        for _ in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            y = e.forward(x)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        row[mode + "_ms"] = round(statistics.median(ts), 4)
        ys[mode] = y
    row["same"] = bool(torch.equal(ys["flow"], ys["phase"]))
    row["flow_over_phase"] = round(row["flow_ms"] / row["phase_ms"], 3)
    out[f"{n}x{h}x{w}"] = row
    del x, ys
print(json.dumps(out))
