#!/usr/bin/env bash
# Builds libvlmfw_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${VLMFW_OUT:-${HERE}/../libvlmfw_b200.so}"
BUILD="${VLMFW_BUILD_DIR:-${HERE}/build}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default
       --expt-relaxed-constexpr -Xptxas -v ${VLMFW_EXTRA_FLAGS:-})
mkdir -p "${BUILD}"
pids=()
for f in gemm_tcgen05 layernorm embed attention attention_tc attention_wb probe probe_deep preprocess sink tower; do
  ( "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${BUILD}/${f}.o" > "${BUILD}/${f}.log" 2>&1 ) &
  pids+=($!)
done
fail=0
for p in "${pids[@]}"; do wait "$p" || fail=1; done
if [ "$fail" -ne 0 ]; then cat "${BUILD}"/*.log >&2; exit 1; fi
"${NVCC}" -shared -o "${OUT}" "${BUILD}"/*.o -lcudart
echo "built ${OUT}"
