#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r6_pytest_tc.log
PFNL_TC_TRACE=1 timeout 120 python tools/tc_trace_test.py fp16x3 2>&1 | tail -9 > gpurun_out/r6_trace_fp16x3.log
for cl in 1; do
  for prec in fp16x3 fp16; do
    PFNL_TC_CLUSTER=$cl timeout 300 python bench.py --steps 10 --warmup 3 --precision $prec --no-cpu-baseline --no-alt > gpurun_out/r6_bench_${prec}_c$cl.json 2> gpurun_out/r6_bench_${prec}_c$cl.err
    echo "cluster=$cl $prec: $(head -c 260 gpurun_out/r6_bench_${prec}_c$cl.json | grep -o '"value": [0-9.e+]*, "unit": "HR-pixels/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*')"; tail -2 gpurun_out/r6_bench_${prec}_c$cl.err
  done
done
grep -E "passed|failed|rror" gpurun_out/r6_pytest_tc.log | tail -5; cat gpurun_out/r6_trace_fp16x3.log | cut -c1-330
