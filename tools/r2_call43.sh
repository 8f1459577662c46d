#!/usr/bin/env bash
# Round-2 GPU call 43: wait for PV(j-1) after the first half of the exp section instead of in front of the turn
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
for v in base pvlate pvlate0; do
  IEF_LIB_PATH=$V/libief_b200_$v.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c43_bench_$v.jsonl 2>&1; echo "$v exit $?"
done
for v in base pvlate pvlate0; do echo "--- $v"; grep -h tcgen05 gpurun_out/r2c43_bench_$v.jsonl | cut -c11-20,128-160; done
IEF_LIB_PATH=$V/libief_b200_pvlate.so timeout 300 python tools/fuzz_attn.py 43 60 2>/dev/null | tail -1
