#!/bin/bash
# (gpurun --gpus 2) the 2-GPU parity test, then both bench arms at N = 2
mkdir -p gpurun_out
T=${1:-r2j}
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -3
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
python - gpurun_out/${T}_bench_2gpu.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('2gpu ms',round(d['ms_per_step'],4),'value %.4e'%d['value'],'e2e %.4e'%d['e2e']['value'],'e2e ms',round(d['e2e']['ms_per_step'],4))
    print('with_collective',{k:v for k,v in d['with_collective'].items() if k!='what'})
except Exception as e:
    print('bench failed',e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -c 300
