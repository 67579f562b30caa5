#!/bin/bash
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r7_pytest_all.log
for prec in fp16x3 fp16; do
  timeout 300 python bench.py --steps 10 --warmup 3 --precision $prec --no-cpu-baseline --no-alt > gpurun_out/r7_bench_$prec.json 2> gpurun_out/r7_bench_$prec.err
done
cat gpurun_out/r7_pytest_all.log
python - <<'PY'
import json
for p in ['fp16x3','fp16']:
    d=json.load(open(f'gpurun_out/r7_bench_{p}.json'))
    print(p,'ms',round(d['ms_per_step'],4),'value %.4e'%d['value'],'e2e %.4e'%d['e2e']['value'],'launches/step',d['launches_per_step'])
    print('  ',{k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})
    r=d['roofline']; print('  roofline',r['kernel'],r['bound'],round(r['achieved'],1),round(r['frac'],3),'other',round(r['other']['achieved'],1),round(r['other']['frac'],3))
PY
