#!/bin/bash
S=$(date +%s)
timeout 120 python __graft_entry__.py smoke 2>&1 | grep -v CUDAEvent | tail -3
echo "smoke took $(( $(date +%s) - S )) s"
S=$(date +%s)
timeout 60 python -u - <<'PY' 2>&1 | grep -v CUDAEvent | tail -4
import sys; sys.path.insert(0,'/root/repo')
import torch
from pfnl_b200 import PFNL, weights as WT
m = PFNL(weights=WT.xavier_init(), device=0, precision='fp16x3')
y = m.forward(torch.rand(1,7,16,16,3,device='cuda'))
raise AssertionError("deliberate failure with a live handle in the traceback")
PY
echo "failing script exited after $(( $(date +%s) - S )) s"
