#!/usr/bin/env bash
# Round-2 GPU call 10: how much of attn_tc3 is the exp pipe? (timing-only variants that drop every n-th exponential) + traces
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
for v in skip4 skip2 skip1; do
  IEF_LIB_PATH=$V/libief_b200_$v.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c10_bench_$v.jsonl 2>&1; echo "$v exit $?"
done
timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c10_bench_default.jsonl 2>&1
grep -h '"sd15_64"\|"big_d40"' gpurun_out/r2c10_bench_*.jsonl | cut -c1-200
