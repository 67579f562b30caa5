#!/bin/bash
for n in 16 8 4; do
echo "== clips $n"
PFNL_TC_TRACE=1 timeout 120 python tools/tc_trace_test.py fp16x3 $n 2>&1 | tail -4 | cut -c1-330
done
