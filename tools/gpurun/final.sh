#!/bin/bash
# Consolidated round measurements on one B200: tests, both bench arms, configs 4/5, ncu launch list and full captures.
mkdir -p gpurun_out
T=${1:-r17}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}_pytest.log
timeout 500 python bench.py > gpurun_out/${T}_bench_fp16x3.json 2> gpurun_out/${T}_bench_fp16x3.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 300 python tools/config_sweep.py > gpurun_out/${T}_config_sweep.json 2> gpurun_out/${T}_config_sweep.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_fp16x3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-alt > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 30 -c 2 -o gpurun_out/${T}_conv_tc -f python tools/dbg_forward.py fp16x3 eager > gpurun_out/${T}_ncu_conv.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:nl_tc -s 7 -c 2 -o gpurun_out/${T}_nl_tc -f python tools/nl_one.py > gpurun_out/${T}_ncu_nl.log 2>&1
cat gpurun_out/${T}_pytest.log
python - $T <<'PY'
import json,sys
T=sys.argv[1]
d=json.load(open(f'gpurun_out/{T}_bench_fp16x3.json'))
print('ms',round(d['ms_per_step'],4),'value %.4e'%d['value'],'e2e %.4e'%d['e2e']['value'],'launches/step',d['launches_per_step'],d['clocks'])
r=d['roofline']; print('  roofline',r['kernel'],r['bound'],round(r['achieved'],1),round(r['frac'],3),'traffic',r['traffic'])
print(d.get('other_precisions')); print(d.get('cpu_baseline'))
print(open(f'gpurun_out/{T}_bench_reference.json').read()[:240])
print(open(f'gpurun_out/{T}_config_sweep.json').read()[:1800])
PY
tail -2 gpurun_out/${T}_ncu_nl.log
