#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r4_pytest_tc.log
PFNL_TC_TRACE=1 timeout 120 python tools/tc_trace_test.py fp16 2>&1 | tail -12 > gpurun_out/r4_trace_fp16.log
PFNL_TC_TRACE=1 timeout 120 python tools/tc_trace_test.py fp16x3 2>&1 | tail -12 > gpurun_out/r4_trace_fp16x3.log
for prec in fp16x3 fp16; do
  timeout 300 python bench.py --steps 10 --warmup 3 --precision $prec --no-cpu-baseline > gpurun_out/r4_bench_$prec.json 2> gpurun_out/r4_bench_$prec.err
done
grep -E "passed|failed|regime|Error|error" gpurun_out/r4_pytest_tc.log | tail -12; cat gpurun_out/r4_trace_fp16.log gpurun_out/r4_trace_fp16x3.log; head -c 300 gpurun_out/r4_bench_fp16x3.json; echo; head -c 300 gpurun_out/r4_bench_fp16.json; echo; tail -3 gpurun_out/r4_bench_fp16x3.err
