#!/bin/bash
# role splits around the throughput-balanced point (conv2f tiles cost 1.3x conv1 tiles)
S="56,10,10,72 56,11,9,72 56,12,8,72 56,13,10,69 53,12,8,75 54,11,8,75 56,11,8,73 52,10,10,76"
for n in 16 32; do echo "== clips $n x 32x32"; SWEEP_N=$n SWEEP_ITERS=12 timeout 600 python tools/flow_split_sweep.py $S; done
echo "== 1 x 180x318"; SWEEP_N=1 SWEEP_H=180 SWEEP_W=318 SWEEP_ITERS=8 timeout 600 python tools/flow_split_sweep.py $S
