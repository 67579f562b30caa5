#!/bin/bash
# ncu: launch list of one bench step + full captures of the PFRB dataflow kernel and of nl_tc_kernel<2> at L = 4096
mkdir -p gpurun_out
T=${1:-ncu}
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -s 330 -c 24 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt --parity-clips 0 --no-nl-roofline > gpurun_out/${T}_launches_bench.log 2>&1
echo "launch list rc=$?"; tail -3 gpurun_out/${T}_launches.csv | cut -c1-200
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:pfrb_flow_kernel -s 1 -c 1 -f -o gpurun_out/${T}_flow python tools/flow_trace.py fp16x3 16 32 > gpurun_out/${T}_flow.log 2>&1
echo "flow capture rc=$?"; tail -2 gpurun_out/${T}_flow.log
ls -la gpurun_out/${T}_*.ncu-rep
