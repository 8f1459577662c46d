#!/usr/bin/env bash
# Round-2 GPU call 41: re-tune the exp-section switches on top of the persistent form
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
for v in default emul3 emul6 scaleturn preturn late; do
  if [ $v = default ]; then L=""; else L="IEF_LIB_PATH=$V/libief_b200_$v.so"; fi
  env $L timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c41_bench_$v.jsonl 2>&1; echo "$v exit $?"
done
for v in default emul3 emul6 scaleturn preturn late; do echo "--- $v"; grep -h tcgen05 gpurun_out/r2c41_bench_$v.jsonl | cut -c11-20,128-160; done
