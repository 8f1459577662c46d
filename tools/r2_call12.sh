#!/usr/bin/env bash
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
for v in trace traceskip; do
  IEF_LIB_PATH=$V/libief_b200_$v.so timeout 120 python tools/tc3_trace.py 4 8 4096 40 > gpurun_out/r2c12_${v}_d40.txt 2>&1; echo "$v d40 exit $?"
  IEF_LIB_PATH=$V/libief_b200_$v.so timeout 120 python tools/tc3_trace.py 4 10 4096 64 > gpurun_out/r2c12_${v}_d64.txt 2>&1; echo "$v d64 exit $?"
done
for f in gpurun_out/r2c12_*.txt; do echo "=== $f"; sed -n '1,1p;6,12p' $f; tail -2 $f; done
