#!/usr/bin/env bash
# Builds go-rio_b200/libapdgicp.so for sm_100a (B200) only.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libapdgicp.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17
       -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xcompiler -Wall -ccbin /usr/bin/g++
       --expt-relaxed-constexpr)
if [[ "${APD_PTXAS_V:-0}" == "1" ]]; then FLAGS+=(-Xptxas -v); fi
if [[ -n "${APD_EXTRA_FLAGS:-}" ]]; then FLAGS+=(${APD_EXTRA_FLAGS}); fi
OUT="${APD_OUT:-${OUT}}"
mkdir -p "${HERE}/_obj"
pids=()
for f in grid knn_cov corr linearize lm prep_ops vgicp apdgicp; do
  "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${HERE}/_obj/${f}.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o "${OUT}" "${HERE}"/_obj/*.o -lpthread -ldl
echo "built ${OUT}"
