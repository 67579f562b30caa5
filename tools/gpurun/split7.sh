#!/bin/bash
# 128 units: does a 12th conv10 CTA pay (12 x 7.0 K = 84.5 K -> 11 x 7.0 K = 77 K cycles per block)?
S="60,11,9,68 61,12,9,66 60,12,9,67 62,12,9,65 64,11,9,64 60,11,9,68"
echo "== clips 16 x 32x32"; SWEEP_N=16 SWEEP_ITERS=20 timeout 600 python tools/flow_split_sweep.py $S
echo "== clips 32 x 32x32"; SWEEP_N=32 SWEEP_ITERS=12 timeout 600 python tools/flow_split_sweep.py 58,11,9,70 59,12,9,68 60,12,9,67 58,12,9,69
