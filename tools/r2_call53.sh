#!/usr/bin/env bash
# Round-2 GPU call 53: K / V ring depth of the persistent 256-row flavour (2 / 3 / 4 stages)
V=image_editing_framework_b200/csrc/build/variants
for v in default st4 st2; do
  if [ $v = default ]; then L=""; else L="IEF_LIB_PATH=$V/libief_b200_$v.so"; fi
  echo "--- $v"; env $L timeout 300 python tools/bench_attn.py tcgen05 big nosdpa 2>/dev/null | grep tcgen05 | cut -c11-20,100-160
done
IEF_LIB_PATH=$V/libief_b200_st4.so timeout 300 python tools/fuzz_attn.py 53 40 2>/dev/null | tail -1
