#!/usr/bin/env bash
# A/B builds of libief_b200.so that differ only in compile-time switches of csrc/attn_tc3.cu.
# Usage: tools/build_variants.sh name1="-DFLAG=1 ..." name2="..."   ->  image_editing_framework_b200/csrc/build/variants/libief_b200_<name>.so
# Select one at run time with IEF_LIB_PATH=<that file>. The default library must have been built first (its other objects are reused).
set -euo pipefail
root="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
src="$root/image_editing_framework_b200/csrc"
out="$src/build/variants"
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr)
pids=()
for spec in "$@"; do
  name="${spec%%=*}"; defs="${spec#*=}"
  read -r -a D <<< "$defs"
  ( "$NVCC" "${FLAGS[@]}" "${D[@]}" -c "$src/attn_tc3.cu" -o "$out/attn_tc3_$name.o"
    "$NVCC" "${FLAGS[@]}" "${D[@]}" -c "$src/api.cu" -o "$out/api_$name.o"   # the trace hook lives in api.cu
    objs=()
    for f in attn_tc attn_tc2 attn_mma cross_attn cross_tc cross_tc_edit cross_attn_bwd elementwise; do objs+=("$src/build/$f.o"); done
    "$NVCC" -shared -o "$out/libief_b200_$name.so" "${objs[@]}" "$out/api_$name.o" "$out/attn_tc3_$name.o" -Xlinker --no-undefined -lcudart_static -lcuda -ldl -lrt -lpthread
    echo "built $out/libief_b200_$name.so" ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
