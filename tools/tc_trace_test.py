import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pfnl_b200 import Engine, weights as WT
prec = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
e = Engine(WT.xavier_init(), 0, prec, graphs=False)
fr = torch.randn(n*7, 32, 32, 64, device='cuda')
for _ in range(2): e.pfrb(3, fr, n, 32, 32)
torch.cuda.synchronize()
