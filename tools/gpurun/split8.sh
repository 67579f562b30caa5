#!/bin/bash
# items rotate over the CTAs of a role across blocks: correctness, then role splits without per-block rounding
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -q -m gpu -x -k "flow or stage or pfrb" 2>&1 | tail -2
S="60,11,9,68 63,12,9,64 64,12,9,63 63,11,9,65 62,12,9,65 64,11,9,64 62,13,9,64 60,11,9,68"
echo "== clips 16 x 32x32"; SWEEP_N=16 SWEEP_ITERS=20 timeout 600 python tools/flow_split_sweep.py $S
echo "== clips 32 x 32x32"; SWEEP_N=32 SWEEP_ITERS=12 timeout 600 python tools/flow_split_sweep.py 58,12,9,69 63,12,9,64 61,12,9,66 60,12,9,67
echo "== 1 x 180x318"; SWEEP_N=1 SWEEP_H=180 SWEEP_W=318 SWEEP_ITERS=8 timeout 600 python tools/flow_split_sweep.py 58,12,9,69 61,12,9,66 63,12,9,64
