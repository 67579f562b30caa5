"""PFNL_TC_TRACE=1 python tools/flow_trace.py [precision] [clips] [size]: per-role timing of the PFRB dataflow kernel
(CTA durations, start/end skew, cycles the producers spent waiting for dependencies) on one forward."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfnl_b200 import Engine, weights as WT

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
size = int(sys.argv[3]) if len(sys.argv) > 3 else 32
e = Engine(WT.xavier_init(), 0, prec, graphs=False)
x = torch.rand(n, 7, size, size, 3, device="cuda")
for _ in range(3):
    e.forward(x)
torch.cuda.synchronize()
