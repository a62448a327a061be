#!/bin/bash
# Build libjc_b200.so (sm_100a only) next to the Python package.  Usage: build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libjc_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
SRCS="jc_plan.cu jc_setup.cu jc_tracers.cu jc_power.cu jc_power_adj.cu jc_contract.cu jc_pipeline.cu jc_gather.cu jc_loglike.cu jc_cl_loglike.cu jc_sparse.cu jc_api.cu"
mkdir -p "$HERE/build"
pids=()
for f in $SRCS; do
  "$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC,-O2,-ffp-contract=off -c "$HERE/$f" -o "$HERE/build/${f%.cu}.o" "$@" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" \
  $(for f in $SRCS; do echo "$HERE/build/${f%.cu}.o"; done) -lcudart
echo "built $OUT"
