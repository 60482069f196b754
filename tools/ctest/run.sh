#!/usr/bin/env bash
# Builds (if needed) and runs the standalone C-ABI check against the in-tree libcsm_b200.so.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
LIBDIR="$HERE/../../csm-train-pytorch_b200"
if [[ ! -x "$HERE/narrow_tail_check" || "$HERE/narrow_tail_check.cu" -nt "$HERE/narrow_tail_check" ]]; then
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --cudart=shared -o "$HERE/narrow_tail_check" \
    "$HERE/narrow_tail_check.cu" -L"$LIBDIR" -lcsm_b200 -Xlinker -rpath -Xlinker "\$ORIGIN/../../csm-train-pytorch_b200"
fi
export LD_LIBRARY_PATH="/usr/local/cuda/lib64:/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/cuda_runtime/lib:${LD_LIBRARY_PATH:-}"
exec "$HERE/narrow_tail_check"
