#!/usr/bin/env bash
# Round-2 GPU call 36: clock64 trace of a persistent CTA (items 0 and 1 per tile, every item start / loop end / epilogue end)
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
for v in trace0 trace1; do
  for shape in "4 8 4096 40" "4 10 4096 64"; do
    echo "=== $v  $shape (pair mode)"
    IEF_TC_SPLITKV=0 IEF_LIB_PATH=$V/libief_b200_$v.so timeout 120 python tools/tc3_trace.py $shape 2>&1 | tail -40
  done
done > gpurun_out/r2c36_trace.txt 2>&1
cat gpurun_out/r2c36_trace.txt
