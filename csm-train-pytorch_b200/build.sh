#!/usr/bin/env bash
# Builds libcsm_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
# A/B builds: CSM_LIB_SUFFIX=_x CSM_EXTRA_NVCC_FLAGS="-DFOO" build.sh -> libcsm_b200_x.so (select it with CSM_B200_LIB)
SUFFIX="${CSM_LIB_SUFFIX:-}"
EXTRA="${CSM_EXTRA_NVCC_FLAGS:-}"
OUT="$HERE/libcsm_b200${SUFFIX}.so"
BUILD="$HERE/build${SUFFIX}"
SRCS=("$HERE"/csrc/*.cu)
rm -f "$BUILD"/stubs_v1.o
mkdir -p "$BUILD"
OBJS=()
pids=()
for s in "${SRCS[@]}"; do
  o="$BUILD/$(basename "${s%.cu}").o"
  OBJS+=("$o")
  if [[ ! -f "$o" || "$s" -nt "$o" || "$HERE/csrc/common.cuh" -nt "$o" || "$HERE/csrc/tc_common.cuh" -nt "$o" || "$HERE/../include/csm_b200.h" -nt "$o" ]]; then
    "$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
      -Xcompiler -fPIC -Xptxas -v $EXTRA -c "$s" -o "$o" > "$BUILD/$(basename "${s%.cu}").log" 2>&1 &
    pids+=($!)
  fi
done
fail=0
for p in "${pids[@]:-}"; do [[ -z "$p" ]] || wait "$p" || fail=1; done
if [[ $fail -ne 0 ]]; then
  for l in "$BUILD"/*.log; do grep -E "error|Error|fatal" "$l" >/dev/null && { echo "== $l"; grep -v "^ptxas info" "$l" | head -40; }; done
  exit 1
fi
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "${OBJS[@]}" -lcudart
echo "built $OUT"
