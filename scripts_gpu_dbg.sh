#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_metrics.py -x -q 2>&1 | grep -v CUDAEvent | tail -30
