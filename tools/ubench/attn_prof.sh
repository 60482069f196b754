#!/usr/bin/env bash
# Builds a PROFILING variant of the library (attn_tc.cu with -DCSM_ATTN_PROF: per-role wait-cycle counters in the attention
# backward kernels) next to the product one and prints the counters for B=2, S=2048 (run on a B200; `build` works anywhere).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
PKG="$HERE/../../csm-train-pytorch_b200"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
if [[ "${1:-}" == "build" || ! -f "$HERE/libcsm_b200_prof.so" ]]; then
  "$PKG/build.sh" > /dev/null
  "$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DCSM_ATTN_PROF ${PROF_DEFS:-} \
    -c "$PKG/csrc/attn_tc.cu" -o "$HERE/attn_tc_prof.o"
  OBJS=()
  for o in "$PKG"/build/*.o; do [[ "$(basename "$o")" == "attn_tc.o" ]] || OBJS+=("$o"); done
  "$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$HERE/libcsm_b200_prof.so" "$HERE/attn_tc_prof.o" "${OBJS[@]}" -lcudart
  echo "built $HERE/libcsm_b200_prof.so"
fi
[[ "${1:-}" == "build" ]] && exit 0
CSM_B200_LIB="$HERE/libcsm_b200_prof.so" python "$HERE/attn_prof.py" "$@"
