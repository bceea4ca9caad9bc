#!/usr/bin/env bash
# Builds libvlmfw_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libvlmfw_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default
       --expt-relaxed-constexpr -Xptxas -v)
mkdir -p "${HERE}/build"
pids=()
for f in gemm_tcgen05 layernorm embed attention attention_tc probe tower; do
  ( "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${HERE}/build/${f}.o" > "${HERE}/build/${f}.log" 2>&1 ) &
  pids+=($!)
done
fail=0
for p in "${pids[@]}"; do wait "$p" || fail=1; done
if [ "$fail" -ne 0 ]; then cat "${HERE}"/build/*.log >&2; exit 1; fi
"${NVCC}" -shared -o "${OUT}" "${HERE}"/build/*.o -lcudart
echo "built ${OUT}"
