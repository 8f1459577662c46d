#!/usr/bin/env bash
# Round-2 GPU call 11: is attn_tc3 at head_dim 40 bound by the NUMBER of tcgen05.mma instructions? (no exponentials; with / without the row-sum MMAs)
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
echo "--- no exponentials, row-sum MMA on (d=40: 19 MMA per stream-tile)"
IEF_LIB_PATH=$V/libief_b200_skip1.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa 2>/dev/null | grep tcgen05 | cut -c1-200 | tee gpurun_out/r2c11_skip1.jsonl
echo "--- no exponentials, row sums in registers (d=40: 11 MMA per stream-tile)"
IEF_TC3_NO_SUM_MMA=1 IEF_LIB_PATH=$V/libief_b200_skip1.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa 2>/dev/null | grep tcgen05 | cut -c1-200 | tee gpurun_out/r2c11_skip1_nosummma.jsonl
echo "--- exponentials on, row sums in registers"
IEF_TC3_NO_SUM_MMA=1 timeout 300 python tools/bench_attn.py tcgen05 big nosdpa 2>/dev/null | grep tcgen05 | cut -c1-200 | tee gpurun_out/r2c11_nosummma.jsonl
