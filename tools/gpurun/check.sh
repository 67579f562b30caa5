#!/bin/bash
# all GPU tests (stop at the first failure), smoke, short benches
mkdir -p gpurun_out
T=${1:-check}
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 200 python __graft_entry__.py smoke 2>&1 | grep -v CUDAEvent | tail -2
for prec in fp16x3; do
  timeout 300 python bench.py --steps 20 --warmup 3 --precision $prec --no-cpu-baseline --no-alt > gpurun_out/${T}_bench_$prec.json 2> gpurun_out/${T}_bench_$prec.err
  python - gpurun_out/${T}_bench_$prec.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print('ms',round(d['ms_per_step'],4),'e2e ms',round(d['e2e']['ms_per_step'],4),{k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})
except Exception as e:
    print('bench failed',e); print(open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
