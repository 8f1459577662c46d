#!/usr/bin/env bash
# Round-2 GPU call 37: persistent 256-row CTAs, lean item loop (even key-tile counts only) — correctness, A/B, trace
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=120 -x -k "tcgen05 or fp16 or row_sources or masactrl or lazy or auto or strided or workspace or full_size or key_bias or probs" > gpurun_out/r2c37_tests.log 2>&1; rc=$?; echo "tests exit $rc"; tail -5 gpurun_out/r2c37_tests.log
[ $rc -ne 0 ] && exit 1
for t in fuzz_attn fuzz_attn_rows; do timeout 300 python tools/$t.py 37 80 2>/dev/null | tail -1; done
IEF_TC_SPLITKV=0 timeout 300 python tools/fuzz_attn.py 38 60 2>/dev/null | tail -1
IEF_TC_SPLITKV=0 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=120 -x -k "tcgen05" 2>&1 | tail -2
timeout 300 python tools/bench_attn.py tcgen05 big > gpurun_out/r2c37_bench_default.jsonl 2>&1; echo "default exit $?"
IEF_TC_SPLITKV=0 timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c37_bench_pair.jsonl 2>&1; echo "pair exit $?"
for v in nopersist nopin; do IEF_LIB_PATH=$V/libief_b200_$v.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c37_bench_$v.jsonl 2>&1; echo "$v exit $?"; done
for v in default pair nopersist nopin; do echo "--- $v"; grep -h "tcgen05\|sdpa" gpurun_out/r2c37_bench_$v.jsonl | cut -c1-60,128-190; done
for shape in "4 8 4096 40" "4 10 4096 64"; do
  echo "=== trace $shape (pair mode)"
  IEF_TC_SPLITKV=0 IEF_LIB_PATH=$V/libief_b200_trace0.so timeout 120 python tools/tc3_trace.py $shape 2>&1 | tail -9
done > gpurun_out/r2c37_trace.txt 2>&1
cat gpurun_out/r2c37_trace.txt
