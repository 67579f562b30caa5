#!/bin/bash
# tools/ab_build.sh <name>: build the current sources into variants/lib_<name>.so (selected with PFNL_B200_LIB) for
# A/B measurements of kernel variants inside ONE gpurun call (different boxes differ by a few per cent).
set -e
cd "$(dirname "$0")/.."
python -m pfnl_b200.build --force > /dev/null
mkdir -p variants
cp pfnl_b200/libpfnl_b200.so variants/lib_$1.so
echo built variants/lib_$1.so
