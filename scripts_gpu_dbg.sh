#!/bin/bash
timeout 300 python tools/gate_margin.py 2>&1 | grep -v CUDAEvent | tail -2
PFNL_NL_FFMA=1 timeout 300 python tools/gate_margin.py 2>&1 | grep -v CUDAEvent | tail -2
