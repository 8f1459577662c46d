#!/usr/bin/env bash
# Round-2 GPU call 42: head_dim 49-64 variant: share of FMA-pipe exponentials x scale inside the exp section
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
for v in e4t e5t e6t e8t e8 e0t; do
  IEF_LIB_PATH=$V/libief_b200_$v.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c42_bench_$v.jsonl 2>&1; echo "$v exit $?"
done
for v in e4t e5t e6t e8t e8 e0t; do echo "--- $v"; grep -h tcgen05 gpurun_out/r2c42_bench_$v.jsonl | grep -v '"d": 40' | cut -c11-20,128-160; done
