#!/bin/bash
mkdir -p gpurun_out
T=${1:-r11}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
timeout 100 python tools/dbg_forward.py fp16x3 profile 2>&1 | grep -v CUDAEvent | tail -2
PFNL_TC_TRACE=1 timeout 120 python tools/tc_trace_test.py fp16x3 2>&1 | tail -12 > gpurun_out/${T}_trace_fp16x3.log
for prec in fp16x3 fp16; do
  timeout 300 python bench.py --steps 20 --warmup 3 --precision $prec --no-cpu-baseline --no-alt > gpurun_out/${T}_bench_$prec.json 2> gpurun_out/${T}_bench_$prec.err
done
python - $T <<'PY'
import json,sys
T=sys.argv[1]
for p in ['fp16x3','fp16']:
    try:
        d=json.load(open(f'gpurun_out/{T}_bench_{p}.json'))
    except Exception as e:
        print(p,'bench failed',e); print(open(f'gpurun_out/{T}_bench_{p}.err').read()[-600:]); continue
    print(p,'ms',round(d['ms_per_step'],4),'value %.4e'%d['value'],'e2e %.4e'%d['e2e']['value'],'launches/step',d['launches_per_step'])
    print('  ',{k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})
PY
cut -c1-330 gpurun_out/${T}_trace_fp16x3.log
