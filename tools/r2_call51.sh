#!/usr/bin/env bash
# Round-2 GPU call 51: active-row bit mask instead of an indexed constant load; where do the ~2 K cycles between two items go at head_dim 40?
V=image_editing_framework_b200/csrc/build/variants
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=120 -x -k "persistent or second_pass or key_bias or row_sources" 2>&1 | tail -2
timeout 300 python tools/bench_attn.py tcgen05 big nosdpa 2>/dev/null | grep tcgen05 | cut -c11-20,100-160
for shape in "4 8 4096 40" "4 10 4096 64"; do echo "=== $shape"; IEF_TC_SPLITKV=0 IEF_LIB_PATH=$V/libief_b200_trace0.so timeout 120 python tools/tc3_trace.py $shape 2>&1 | tail -6; done
