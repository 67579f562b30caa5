#!/bin/bash
# Last check of a build: every GPU test, smoke(), the default bench line.
mkdir -p gpurun_out
T=${1:-verify}
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py smoke 2>&1 | grep -v CUDAEvent | tail -2
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - $T <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/{sys.argv[1]}_bench.json'))
print('ms',round(d['ms_per_step'],4),'value %.4e'%d['value'],'e2e %.4e'%d['e2e']['value'],'launches',d['gpu_launches'],d['clocks'])
print(d['roofline']['kernel'],d['roofline']['bound'],round(d['roofline']['frac'],3),d['cpu_baseline']['value'])
PY
