#!/bin/bash
# conv1 : conv2f balance of the dataflow kernel
S="64,10,10,64 60,10,10,68 56,10,10,72 68,10,10,60 52,10,10,76 60,12,10,66"
for n in 32 16; do echo "== clips $n x 32x32"; SWEEP_N=$n SWEEP_ITERS=15 timeout 600 python tools/flow_split_sweep.py $S; done
