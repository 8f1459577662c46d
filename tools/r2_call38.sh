#!/usr/bin/env bash
# Round-2 GPU call 38: why are the split-KV CTAs of the remainder launch slow? trace + pure split timing
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
for shape in "4 8 4096 40" "4 10 4096 64"; do
  echo "=== trace $shape (split mode)"
  IEF_TC_SPLITKV=1 IEF_LIB_PATH=$V/libief_b200_trace0.so timeout 120 python tools/tc3_trace.py $shape 2>&1 | tail -22
done > gpurun_out/r2c38_trace.txt 2>&1
cat gpurun_out/r2c38_trace.txt
IEF_TC_SPLITKV=1 timeout 300 python tools/bench_attn.py tcgen05 big nosdpa 2>&1 | grep tcgen05 | cut -c1-60,128-190
