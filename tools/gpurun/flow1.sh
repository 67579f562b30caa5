#!/bin/bash
# check of the PFRB dataflow kernel: debug cases, bit-exactness vs the phase kernels, trace, short bench, full suite
mkdir -p gpurun_out
T=${1:-flow1}
timeout 400 python tools/flow_debug.py 2>&1 | cut -c1-200 | head -60 > gpurun_out/${T}_debug.log
cat gpurun_out/${T}_debug.log | head -20
if grep -q "FAILED\|TIMEOUT\|MISMATCH" gpurun_out/${T}_debug.log; then echo "FLOW DEBUG CASES FAILED"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_tensorcore.py -x -q -k "flow" 2>&1 | tail -15 > gpurun_out/${T}_flowtests.log
cat gpurun_out/${T}_flowtests.log
if grep -q "failed\|error\|Error" gpurun_out/${T}_flowtests.log; then echo "FLOW TESTS FAILED"; exit 1; fi
PFNL_TC_TRACE=1 timeout 120 python tools/flow_trace.py fp16x3 16 32 2>&1 | grep -v CUDAEvent | tail -12 > gpurun_out/${T}_flow_trace.log
cat gpurun_out/${T}_flow_trace.log
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
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
