#!/usr/bin/env bash
# Build liblfe.so for sm_100a (B200) in-tree.  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../lib"
mkdir -p "${OUT}" "${HERE}/build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -O3
       --expt-relaxed-constexpr -Xptxas -v)
SRCS=(api prims sort factorize demean gram vcov compress shard upload)
pids=()
for s in "${SRCS[@]}"; do
  if [ ! -f "${HERE}/build/${s}.o" ] || [ "${HERE}/${s}.cu" -nt "${HERE}/build/${s}.o" ] || \
     [ -n "$(find "${HERE}" -maxdepth 1 -name '*.cuh' -newer "${HERE}/build/${s}.o" 2>/dev/null)" ] || \
     [ "${HERE}/../../include/lfe.h" -nt "${HERE}/build/${s}.o" ]; then
    ( "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${s}.cu" -o "${HERE}/build/${s}.o" > "${HERE}/build/${s}.log" 2>&1 \
        || { cat "${HERE}/build/${s}.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
OBJS=()
for s in "${SRCS[@]}"; do OBJS+=("${HERE}/build/${s}.o"); done
"${NVCC}" -shared -o "${OUT}/liblfe.so" "${OBJS[@]}" -lcudart -ldl -lpthread
echo "built ${OUT}/liblfe.so"
