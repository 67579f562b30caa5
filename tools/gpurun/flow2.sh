#!/bin/bash
# iteration check of the PFRB dataflow kernel: debug cases (bit-exactness), trace, short benches with role splits
mkdir -p gpurun_out
T=${1:-flow}
timeout 300 python tools/flow_debug.py 2>&1 | cut -c1-200 | head -40 > gpurun_out/${T}_debug.log
cat gpurun_out/${T}_debug.log | head -12
if grep -q "FAILED\|TIMEOUT\|MISMATCH" gpurun_out/${T}_debug.log; then echo "FLOW DEBUG CASES FAILED"; exit 1; fi
PFNL_TC_TRACE=1 timeout 120 python tools/flow_trace.py fp16x3 16 32 2>&1 | grep -v CUDAEvent | grep -A40 "flow-trace" | tail -22 | cut -c1-1500 > gpurun_out/${T}_trace.log
grep "x *[0-9]*:" gpurun_out/${T}_trace.log | cut -c1-330
for split in default "64,11,9,64"; do
  if [ "$split" != default ]; then export PFNL_FLOW_SPLIT=$split; fi
  timeout 200 python bench.py --steps 20 --warmup 3 --precision fp16x3 --no-cpu-baseline --no-alt > gpurun_out/${T}_bench_${split//,/_}.json 2> gpurun_out/${T}_bench_${split//,/_}.err
  python - "$split" gpurun_out/${T}_bench_${split//,/_}.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print('split',sys.argv[1],'ms',round(d['ms_per_step'],4),'e2e ms',round(d['e2e']['ms_per_step'],4),{k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})
except Exception as e:
    print('split',sys.argv[1],'bench failed',e)
PY
done
