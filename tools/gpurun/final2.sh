#!/bin/bash
# final verification of the round: GPU tests, smoke, bench (both arms), memcheck over the dataflow kernel
mkdir -p gpurun_out
T=${1:-r2f}
tools/gpurun/check2.sh $T
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 500; echo
export PFNL_TC_WAIT_LIMIT_CYCLES=2000000000000
timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_forward.py fp16x3 > gpurun_out/${T}_memcheck.log 2>&1
echo "memcheck rc=$?: $(grep -E 'ERROR SUMMARY' gpurun_out/${T}_memcheck.log | tail -1) $(grep -c ' ok ' gpurun_out/${T}_memcheck.log) ok"
