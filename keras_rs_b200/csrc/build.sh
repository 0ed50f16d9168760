#!/bin/bash
# Builds libkrs_b200.so for sm_100a, in-tree.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v --expt-relaxed-constexpr)
# e.g. KRS_EXTRA_FLAGS=-DKRS_TC_TRACE=1 (per-role clock64 timeline of gemm_tc_kernel, tests/tc_trace.py); touch the source first
if [ -n "${KRS_EXTRA_FLAGS:-}" ]; then FLAGS+=(${KRS_EXTRA_FLAGS}); fi
SRCS=(api gemm_ffma gemm_tc cross_dense gather optim dot topk shard exchange rowops)
pids=()
for s in "${SRCS[@]}"; do
  if [ ! -f "$HERE/obj/$s.o" ] || [ "$HERE/$s.cu" -nt "$HERE/obj/$s.o" ] || [ "$HERE/common.cuh" -nt "$HERE/obj/$s.o" ] || [ "$HERE/tc_common.cuh" -nt "$HERE/obj/$s.o" ] || [ "$HERE/../../include/krs_b200.h" -nt "$HERE/obj/$s.o" ]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$HERE/$s.cu" -o "$HERE/obj/$s.o" > "$HERE/obj/$s.log" 2>&1 || { cat "$HERE/obj/$s.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
OBJS=()
for s in "${SRCS[@]}"; do OBJS+=("$HERE/obj/$s.o"); done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libkrs_b200.so" "${OBJS[@]}" -lcudart
echo "built $OUT/libkrs_b200.so"
