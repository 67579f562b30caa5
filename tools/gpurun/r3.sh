#!/bin/bash
# Third GPU pass: rewritten TC conv issue loop + PDL + accumulation chains, tcgen05 non-local.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r3_pytest_tc.log
PFNL_TC_CHAINS=1 timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -s -k "pr1_gate or small_fp16x3" 2>&1 | tail -12 > gpurun_out/r3_pytest_chains1.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "pr1_parity" 2>&1 | tail -8 > gpurun_out/r3_pytest_fp32_gate.log
for prec in fp16x3 fp16; do
  timeout 300 python bench.py --steps 10 --warmup 3 --precision $prec --no-cpu-baseline > gpurun_out/r3_bench_$prec.json 2> gpurun_out/r3_bench_$prec.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --precision fp16x3 --no-graphs --no-cpu-baseline > gpurun_out/r3_bench_fp16x3_nograph.json 2>> gpurun_out/r3_bench_fp16x3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 12 -c 4 -o gpurun_out/r3_conv_tc_fp16x3 python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu-baseline --precision fp16x3 > gpurun_out/r3_ncu_full.log 2>&1
tail -25 gpurun_out/r3_pytest_tc.log; cat gpurun_out/r3_pytest_chains1.log gpurun_out/r3_pytest_fp32_gate.log | grep -E "regime|passed|failed"; head -c 400 gpurun_out/r3_bench_fp16x3.json; echo; tail -3 gpurun_out/r3_bench_fp16x3.err
