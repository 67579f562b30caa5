#!/bin/bash
timeout 100 python tools/dbg_pfrb.py fp16x3 1 32 32 2>&1 | grep -v CUDAEvent | tail -5
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -x -q -s -k "test_single_pfrb" 2>&1 | grep -v CUDAEvent | tail -25
