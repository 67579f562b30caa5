#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -x -q -k "nonlocal or forward_128 or 6480" 2>&1 | grep -v CUDAEvent | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nl_ --csv --log-file gpurun_out/nl_launches.csv python tools/nl_one.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/nl_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value'); gi=H.index('Grid Size')
for r in rows[hdr+1:]:
    if len(r)>vi: print(r[ki][:40], r[gi], r[vi])
PY
timeout 300 python bench.py --steps 20 --warmup 3 --precision fp16x3 --no-cpu-baseline --no-alt 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fp16x3 ms',d['ms_per_step'],'value %.4e'%d['value'],{k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
