#!/bin/bash
# Build libjc_b200.so (sm_100a only) next to the Python package.  Usage: build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libjc_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC,-O2,-ffp-contract=off -shared \
  -o "$OUT" "$HERE/jc_plan.cu" "$HERE/jc_pipeline.cu" "$HERE/jc_api.cu" -lcudart "$@"
echo "built $OUT"
