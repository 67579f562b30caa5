#!/bin/bash
# Second GPU pass: tensor-core conv path parity + benches + ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -s 2>&1 | tail -80 > gpurun_out/r2_pytest_tc.log
for prec in fp16x3 fp16; do
  timeout 300 python bench.py --steps 10 --warmup 3 --precision $prec --no-cpu-baseline > gpurun_out/r2_bench_$prec.json 2> gpurun_out/r2_bench_$prec.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_fp16x3.csv python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu-baseline --precision fp16x3 > gpurun_out/r2_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 12 -c 4 -o gpurun_out/r2_conv_tc_fp16x3 python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu-baseline --precision fp16x3 > gpurun_out/r2_ncu_full.log 2>&1
tail -30 gpurun_out/r2_pytest_tc.log; head -c 1500 gpurun_out/r2_bench_fp16x3.json; echo; tail -3 gpurun_out/r2_bench_fp16x3.err
