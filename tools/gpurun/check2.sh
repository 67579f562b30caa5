#!/bin/bash
# GPU tests with printed error levels for the gate tests, then bench
mkdir -p gpurun_out
T=${1:-check}
timeout 1500 python -m pytest tests -m gpu -q -rA 2>&1 | grep -E "passed|failed|max-abs|vs fp64|FAILED|Error" | cut -c1-220 | tail -70 > gpurun_out/${T}_pytest.log
tail -25 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 3 --precision fp16x3 --no-alt --cpu-seconds 8 > gpurun_out/${T}_bench_fp16x3.json 2> gpurun_out/${T}_bench_fp16x3.err
python - gpurun_out/${T}_bench_fp16x3.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print('ms',round(d['ms_per_step'],4),'e2e ms',round(d['e2e']['ms_per_step'],4),'blocking',round(d['e2e']['blocking_call_ms_per_step'],4),'f64 pageable',round(d['e2e']['pageable_float64_input']['ms_per_step'],4),round(d['e2e']['pageable_float64_input']['blocking_call_ms_per_step'],4))
    print('collective',d['with_collective']['ms_per_step'],d['with_collective']['collective_us'])
    print('parity',d['parity_check'])
    print('nl',d['nonlocal_roofline'])
    print('roofline',{k:v for k,v in d['roofline'].items() if k in('kernel','bound','achieved','frac','avg_launch_ms','share_of_step')},d['roofline']['other'])
    print({k:round(v,3) for k,v in d['kernel_ms_per_step'].items()}, d['cpu_baseline'])
except Exception as e:
    print('bench failed',e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
