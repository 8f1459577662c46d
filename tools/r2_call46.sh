#!/usr/bin/env bash
# Round-2 GPU call 46: hand the turn over on entering the exp section; one tcgen05.ld.x64 per tile
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
for v in default early ld64; do
  if [ $v = default ]; then L=""; else L="IEF_LIB_PATH=$V/libief_b200_$v.so"; fi
  env $L timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c46_bench_$v.jsonl 2>&1; echo "$v exit $?"
done
for v in default early ld64; do echo "--- $v"; grep -h tcgen05 gpurun_out/r2c46_bench_$v.jsonl | cut -c11-20,128-160; done
for v in early ld64; do IEF_LIB_PATH=$V/libief_b200_$v.so timeout 300 python tools/fuzz_attn.py 46 40 2>/dev/null | tail -1; done
