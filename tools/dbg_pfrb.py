"""Debug helper: one PFRB through the fp32-io wrapper (f32->planes, 2 TC launches, planes->f32)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from pfnl_b200 import Engine, weights as WT
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp16x3'
n, h, w = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (1, 32, 32)
e = Engine(WT.xavier_init(), 0, prec, graphs=False)
fr = torch.randn(n * 7, h, w, 64, device='cuda')
out = e.pfrb(3, fr, n, h, w)
torch.cuda.synchronize()
print('pfrb ok', float(out.abs().max()))
