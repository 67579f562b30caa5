#!/bin/bash
# Consolidation pass: full GPU suite, smoke, headline bench + reference arm, ncu for the final kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r5_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r5_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r5_bench_default.json 2> gpurun_out/r5_bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r5_bench_reference.json 2>> gpurun_out/r5_bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r5_launches_fp16x3.csv python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu-baseline --no-alt > gpurun_out/r5_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 12 -c 4 -o gpurun_out/r5_conv_tc_fp16x3 python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu-baseline --no-alt > gpurun_out/r5_ncu_full.log 2>&1
cat gpurun_out/r5_pytest_all.log; tail -2 gpurun_out/r5_smoke.log; head -c 600 gpurun_out/r5_bench_default.json; echo; tail -3 gpurun_out/r5_bench_default.err
