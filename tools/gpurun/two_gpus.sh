#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_gpu_check.py 2>&1 | grep -v CUDAEvent | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r18_bench_2gpu.json 2> gpurun_out/r18_bench_2gpu.err
tail -c 1500 gpurun_out/r18_bench_2gpu.json | head -c 1500; echo
python -c "
import json; d=json.loads(open('gpurun_out/r18_bench_2gpu.json').read().strip().splitlines()[-1]); print('2gpu ms',d['ms_per_step'],'value %.4e'%d['value'],'e2e %.4e'%d['e2e']['value'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -c 400
