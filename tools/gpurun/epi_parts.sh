#!/bin/bash
# free-running roles (PFNL_FLOW_DBG bit 2) with parts of the epilogue left out: 4 partial-sum loads, 8 residual loads,
# 16 stores.  Timing only.
for d in 2 6 10 18 30; do
  echo "== PFNL_FLOW_DBG=$d"
  PFNL_FLOW_DBG=$d PFNL_TC_TRACE=1 timeout 300 python tools/flow_trace.py fp16x3 ${1:-16} 32 2>&1 | grep -E " x ?[ 0-9]+: CTA" | tail -4 | cut -c1-110
done
