#!/bin/bash
# compute-sanitizer memcheck / racecheck / synccheck over tools/sanitize_forward.py.
# The PFRB dataflow kernel needs all of its CTAs resident and making progress together; if a tool serialises CTAs it
# cannot run under it, so every tool is tried with the dataflow kernel first and, if that run does not finish
# cleanly, again with PFNL_TC_FLOW=0 (the phase kernels: same tile code, kernel-boundary ordering).
mkdir -p gpurun_out
T=${1:-san}
export PFNL_TC_WAIT_LIMIT_CYCLES=2000000000000
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck; do
  for flow in 1 0; do
    log=gpurun_out/${T}_${tool}_flow${flow}.log
    PFNL_TC_FLOW=$flow timeout 900 $CS --tool $tool --print-limit 20 python tools/sanitize_forward.py fp16x3 fp16 > $log 2>&1
    rc=$?
    echo "== $tool flow=$flow rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) $(grep -c ' ok ' $log) precisions ok"
    if [ $rc -eq 0 ] && grep -q "fp16 ok" $log; then break; fi
  done
done
log=gpurun_out/${T}_memcheck_fp32.log
timeout 600 $CS --tool memcheck --print-limit 20 python tools/sanitize_forward.py fp32 > $log 2>&1
echo "== memcheck fp32 rc=$?: $(grep -E 'ERROR SUMMARY' $log | tail -1)"
for f in gpurun_out/${T}_*.log; do echo "--- $f"; grep -E "=========|ok" $f | grep -v "^========= *$" | cut -c1-200 | tail -12; done
