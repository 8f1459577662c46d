#!/usr/bin/env bash
# Round-2 GPU call 45: item coordinates through host-made reciprocals instead of integer division — correctness, timing, trace
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=120 -x -k "tcgen05 or fp16 or row_sources or masactrl or lazy or auto or strided or workspace or full_size or key_bias or probs" 2>&1 | tail -2
for t in fuzz_attn fuzz_attn_rows; do timeout 300 python tools/$t.py 45 60 2>/dev/null | tail -1; done
timeout 300 python tools/bench_attn.py tcgen05 big nosdpa 2>/dev/null | grep tcgen05 | cut -c11-20,128-160
for shape in "4 8 4096 40" "4 20 1024 64"; do
  echo "=== trace $shape (pair mode)"
  IEF_TC_SPLITKV=0 IEF_LIB_PATH=$V/libief_b200_trace0.so timeout 120 python tools/tc3_trace.py $shape 2>&1 | tail -7
done
