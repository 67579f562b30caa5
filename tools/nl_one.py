"""python tools/nl_one.py [precision] [clips] [L]: three calls of the non-local block in isolation (ncu target)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfnl_b200 import Engine, weights as WT

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
e = Engine(WT.xavier_init(), 0, prec, graphs=False)
t = torch.rand(n, L, 84, device='cuda')
for _ in range(3):
    e.nonlocal_block(t)
torch.cuda.synchronize()
print("nl_one", prec, n, L, "ok")
