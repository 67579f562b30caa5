#!/bin/bash
# after the 256-bit partial-sum loads: free-running tile costs, then role splits
PFNL_FLOW_DBG=2 PFNL_TC_TRACE=1 timeout 300 python tools/flow_trace.py fp16x3 16 32 2>&1 | grep -E " x ?[ 0-9]+: CTA" | tail -4 | cut -c1-110
S="64,10,10,64 61,11,9,67 60,11,9,68 59,12,9,68 58,11,9,70 56,11,9,72 54,11,8,75"
for n in 16 32; do echo "== clips $n x 32x32"; SWEEP_N=$n SWEEP_ITERS=12 timeout 600 python tools/flow_split_sweep.py $S; done
echo "== 1 x 180x318"; SWEEP_N=1 SWEEP_H=180 SWEEP_W=318 SWEEP_ITERS=8 timeout 600 python tools/flow_split_sweep.py 60,11,9,68 58,11,9,70 56,11,9,72 54,11,8,75
