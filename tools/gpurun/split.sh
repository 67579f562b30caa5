#!/bin/bash
# role splits of the dataflow kernel at 128 / 256 / 512 units and at a 180x318 frame
S="64,10,10,64 63,12,9,64 62,13,10,63 62,14,9,63 61,14,10,63 60,16,10,62 60,15,11,62 58,18,12,60"
for n in 16 32 64; do echo "== clips $n x 32x32"; SWEEP_N=$n SWEEP_ITERS=15 timeout 600 python tools/flow_split_sweep.py $S; done
echo "== 1 x 180x318"; SWEEP_N=1 SWEEP_H=180 SWEEP_W=318 SWEEP_ITERS=10 timeout 600 python tools/flow_split_sweep.py $S
