#!/bin/bash
# role splits of the dataflow kernel: split.sh <clips> <H> <W> <split> [<split> ...]   (split = conv1,conv10,conv2b,conv2f)
# All candidates of one comparison belong in ONE gpurun call: boxes differ by a few per cent.
n=${1:-16}; h=${2:-32}; w=${3:-32}; shift 3
echo "== clips $n x ${h}x${w}"
SWEEP_N=$n SWEEP_H=$h SWEEP_W=$w SWEEP_ITERS=${SWEEP_ITERS:-15} timeout 900 python tools/flow_split_sweep.py "$@"
