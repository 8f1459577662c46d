#!/usr/bin/env bash
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
for v in noscale noscale3; do
  IEF_LIB_PATH=$V/libief_b200_$v.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c20_bench_$v.jsonl 2>&1; echo "$v exit $?"
done
for v in noscale noscale3; do echo "--- $v"; grep -h tcgen05 gpurun_out/r2c20_bench_$v.jsonl | cut -c1-60,128-190; done
IEF_LIB_PATH=$V/libief_b200_trace.so timeout 120 python tools/tc3_trace.py 4 8 4096 40 > gpurun_out/r2c20_trace_d40.txt 2>&1; sed -n '1,1p;6,9p' gpurun_out/r2c20_trace_d40.txt; tail -2 gpurun_out/r2c20_trace_d40.txt
IEF_LIB_PATH=$V/libief_b200_trace.so timeout 120 python tools/tc3_trace.py 4 20 1024 64 > gpurun_out/r2c20_trace_sdxl32.txt 2>&1; sed -n '3,6p' gpurun_out/r2c20_trace_sdxl32.txt; tail -2 gpurun_out/r2c20_trace_sdxl32.txt
