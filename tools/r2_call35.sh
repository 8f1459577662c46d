#!/usr/bin/env bash
# Round-2 GPU call 35: persistent 256-row CTAs (IEF_TC3_PERSIST) — correctness, then A/B against the one-item form
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=120 -x -k "tcgen05 or fp16 or row_sources or masactrl or lazy or auto or strided or workspace or full_size or key_bias or probs" > gpurun_out/r2c35_tests.log 2>&1; rc=$?; echo "tests exit $rc"; tail -5 gpurun_out/r2c35_tests.log
if [ $rc -ne 0 ]; then
  echo "--- pure pair mode"; IEF_TC_SPLITKV=0 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=120 -x -k "tcgen05" 2>&1 | tail -5
  exit 1
fi
for t in fuzz_attn fuzz_attn_rows; do timeout 300 python tools/$t.py 35 80 2>/dev/null | tail -1; done
IEF_TC_SPLITKV=0 timeout 300 python tools/fuzz_attn.py 36 60 2>/dev/null | tail -1
IEF_TC_SPLITKV=0 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=120 -x -k "tcgen05" 2>&1 | tail -2
timeout 300 python tools/bench_attn.py tcgen05 big > gpurun_out/r2c35_bench_default.jsonl 2>&1; echo "default exit $?"
IEF_TC_SPLITKV=0 timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c35_bench_pair.jsonl 2>&1; echo "pair exit $?"
IEF_LIB_PATH=$V/libief_b200_nopersist.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c35_bench_nopersist.jsonl 2>&1; echo "nopersist exit $?"
for v in default pair nopersist; do echo "--- $v"; grep -h "tcgen05\|sdpa" gpurun_out/r2c35_bench_$v.jsonl | cut -c1-60,128-190; done
