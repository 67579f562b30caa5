#!/bin/bash
# Re-entry check of HEAD on a fresh B200: all GPU tests, default bench (both arms), launch list.
mkdir -p gpurun_out
S=$(date +%s)
timeout 1000 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r8_pytest_all.log
echo "pytest_s $(( $(date +%s) - S ))" >> gpurun_out/r8_pytest_all.log
timeout 400 python bench.py > gpurun_out/r8_bench_fp16x3.json 2> gpurun_out/r8_bench_fp16x3.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r8_bench_reference.json 2> gpurun_out/r8_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r8_launches_fp16x3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-alt > gpurun_out/r8_ncu_bench.log 2>&1
PFNL_TC_TRACE=1 timeout 120 python tools/tc_trace_test.py fp16x3 2>&1 | tail -12 > gpurun_out/r8_trace_fp16x3.log
cat gpurun_out/r8_pytest_all.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r8_bench_fp16x3.json'))
print('ms',round(d['ms_per_step'],4),'value %.4e'%d['value'],'e2e %.4e'%d['e2e']['value'],'launches/step',d['launches_per_step'])
print('  ',{k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})
r=d['roofline']; print('  roofline',r['kernel'],r['bound'],round(r['achieved'],1),round(r['frac'],3))
print(d.get('other_precisions'))
print(open('gpurun_out/r8_bench_reference.json').read()[:300])
PY
cut -c1-300 gpurun_out/r8_trace_fp16x3.log
