#!/bin/bash
# the dataflow kernel at 256 units (32 clips of 32x32): trace + ncu full capture, to compare with the 128-unit captures
mkdir -p gpurun_out
T=${1:-r2y}
NCU=/usr/local/cuda/bin/ncu
PFNL_TC_TRACE=1 timeout 300 python tools/flow_trace.py fp16x3 32 32 > gpurun_out/${T}_trace_256u.txt 2>&1
echo "trace rc=$?"; grep -E "x (64|10):|span" gpurun_out/${T}_trace_256u.txt | cut -c1-260
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:pfrb_flow_kernel -s 1 -c 1 -f -o gpurun_out/${T}_flow_256u python tools/flow_trace.py fp16x3 32 32 > gpurun_out/${T}_flow_256u.log 2>&1
echo "flow capture rc=$?"; tail -2 gpurun_out/${T}_flow_256u.log
