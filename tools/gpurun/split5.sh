#!/bin/bash
# 3 k-steps for the 48-channel sources: correctness of the dataflow kernel, free-running tile costs, role splits
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -q -m gpu -x -k "flow or stage or pfrb" 2>&1 | tail -3
PFNL_FLOW_DBG=2 PFNL_TC_TRACE=1 timeout 300 python tools/flow_trace.py fp16x3 16 32 2>&1 | grep -E " x ?[ 0-9]+: CTA" | tail -4 | cut -c1-110
S="60,11,9,68 64,11,9,64 66,11,9,62 68,11,9,60 69,11,8,60 70,11,8,59 72,11,8,57 75,10,7,56"
for n in 16 32; do echo "== clips $n x 32x32"; SWEEP_N=$n SWEEP_ITERS=12 timeout 600 python tools/flow_split_sweep.py $S; done
echo "== 1 x 180x318"; SWEEP_N=1 SWEEP_H=180 SWEEP_W=318 SWEEP_ITERS=8 timeout 600 python tools/flow_split_sweep.py 64,11,9,64 68,11,9,60 70,11,8,59 72,11,8,57
