import sys; sys.path.insert(0,'/root/repo')
import torch
from pfnl_b200 import Engine, weights as WT
e = Engine(WT.xavier_init(), 0, 'fp16', graphs=False)
for (n,L) in [(64,256),(16,1024),(4,4096)]:
    t = torch.rand(n, L, 84, device='cuda')
    for _ in range(3): e.nonlocal_block(t)
    torch.cuda.synchronize()
