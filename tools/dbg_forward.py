"""Debug helper: eager forwards at the bench shape with/without profiling, sync + error check after each."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from pfnl_b200 import Engine, weights as WT
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp16x3'
mode = sys.argv[2] if len(sys.argv) > 2 else 'profile'
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
e = Engine(WT.xavier_init(), 0, prec, graphs=(mode == 'graphs'))
x = torch.rand(n, 7, 32, 32, 3, device='cuda')
out = torch.empty(n, 1, 128, 128, 3, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
if mode == 'profile':
    e.profile(True)
for i in range(6):
    flush.zero_()
    e.forward(x, out=out)
    torch.cuda.synchronize()
    print(prec, mode, 'forward', i, 'ok', float(out.abs().max()), flush=True)
