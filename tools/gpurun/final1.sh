#!/bin/bash
# video sizes (dataflow vs phase kernels), crossover table, then the ncu launch list + full capture of the dataflow kernel
mkdir -p gpurun_out
T=${1:-r2zz}
timeout 600 python tools/video_size_check.py > gpurun_out/${T}_video_sizes.json 2> gpurun_out/${T}_video_sizes.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_video_sizes.json'))
for k,v in d.items():
    print(k,{p:(round(r['ms_per_frame'],3),'%.1e'%r['max_abs_vs_fp32']) for p,r in v.items()})"
timeout 300 python tools/flow_crossover.py > gpurun_out/${T}_crossover.json 2>> gpurun_out/${T}_video_sizes.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_crossover.json'))
for k,v in d.items(): print('  ',k,v)"
tools/gpurun/ncu.sh ${T}
