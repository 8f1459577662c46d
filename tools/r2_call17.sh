#!/usr/bin/env bash
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
for v in pre4 pre3 pre2; do
  IEF_LIB_PATH=$V/libief_b200_$v.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c17_bench_$v.jsonl 2>&1; echo "$v exit $?"
done
timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c17_bench_default.jsonl 2>&1
for v in default pre4 pre3 pre2; do echo "--- $v"; grep -h tcgen05 gpurun_out/r2c17_bench_$v.jsonl | cut -c1-60,128-190; done
