#!/bin/bash
# First GPU pass: parity tests, tcgen05 probe, fp32 bench, ncu launch list + one full capture.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/r1_env.log 2>&1; lscpu | head -20 >> gpurun_out/r1_env.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r1_pytest.log
timeout 120 ./probes/bin/umma_probe > gpurun_out/r1_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/r1_probe.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench_fp32.json 2> gpurun_out/r1_bench_fp32.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/r1_bench_fp32_nograph.json 2>> gpurun_out/r1_bench_fp32.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r1_bench_reference.json 2>> gpurun_out/r1_bench_fp32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1_launches_fp32.csv python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/r1_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ffma_kernel -s 6 -c 2 -o gpurun_out/r1_conv_ffma python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1
tail -5 gpurun_out/r1_pytest.log; cat gpurun_out/r1_probe.log; cat gpurun_out/r1_bench_fp32.json
