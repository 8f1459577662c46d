#!/usr/bin/env bash
# Round-2 GPU call 40: fine phase trace (TRACE=2) of the persistent kernel: where does a stream wait between two exp sections?
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
for shape in "4 8 4096 40" "4 10 4096 64"; do
  echo "=== trace2 $shape (pair mode)"
  IEF_TC_SPLITKV=0 IEF_LIB_PATH=$V/libief_b200_trace2.so timeout 120 python tools/tc3_trace.py $shape 2>&1 | tail -26
done > gpurun_out/r2c40_trace.txt 2>&1
cat gpurun_out/r2c40_trace.txt
