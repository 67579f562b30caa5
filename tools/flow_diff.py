"""Where does the dataflow kernel differ from the phase kernels?  One PFRB, fixed seed; prints the pattern of
differing elements (frame, tile row/col, channel) and whether the dataflow result repeats."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfnl_b200 import Engine, weights as WT

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
n, h, w = 2, 32, 32
e = Engine(WT.xavier_init(), 0, prec, graphs=False)
g = torch.Generator().manual_seed(5)
fr = torch.randn(n * 7, h, w, 64, generator=g).cuda()
e.set_flow(False)
a = e.pfrb(3, fr, n, h, w).clone()
a2 = e.pfrb(3, fr, n, h, w).clone()
e.set_flow(True)
b = e.pfrb(3, fr, n, h, w).clone()
b2 = e.pfrb(3, fr, n, h, w).clone()
torch.cuda.synchronize()
print(prec, "phase repeats:", torch.equal(a, a2), "flow repeats:", torch.equal(b, b2), "flow==phase:", torch.equal(a, b))
d = (a - b).abs()
idx = d.nonzero()
print("differing elements:", idx.shape[0], "of", d.numel(), "max", float(d.max()))
if idx.shape[0]:
    import collections
    print("by frame image:", sorted(collections.Counter((idx[:, 0] % 7).tolist()).items()))
    print("by clip:", sorted(collections.Counter((idx[:, 0] // 7).tolist()).items()))
    print("by tile (y//16, x//8):", sorted(collections.Counter(zip((idx[:, 1] // 16).tolist(), (idx[:, 2] // 8).tolist())).items()))
    print("by channel//16:", sorted(collections.Counter((idx[:, 3] // 16).tolist()).items()))
    print("by row in tile:", sorted(collections.Counter((idx[:, 1] % 16).tolist()).items()))
