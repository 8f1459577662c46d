#!/usr/bin/env bash
# Builds libief_b200.so (sm_100a only) in-tree. Usage: build.sh [out_dir]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${1:-$here/..}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr)
# experiment switches, e.g. IEF_EXTRA_NVCC_FLAGS="-DIEF_TC3_FINE_TRACE=1"
read -r -a EXTRA <<< "${IEF_EXTRA_NVCC_FLAGS:-}"
FLAGS+=("${EXTRA[@]}")
objs=()
pids=()
mkdir -p "$here/build"
for f in api attn_tc attn_tc2 attn_tc3 attn_mma cross_attn cross_tc cross_tc_edit cross_attn_bwd elementwise; do
  "$NVCC" "${FLAGS[@]}" -c "$here/$f.cu" -o "$here/build/$f.o" &
  pids+=($!)
  objs+=("$here/build/$f.o")
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -o "$out/libief_b200.so" "${objs[@]}" -Xlinker --no-undefined -lcudart_static -lcuda -ldl -lrt -lpthread
echo "built $out/libief_b200.so"
