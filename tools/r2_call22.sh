#!/usr/bin/env bash
# Round-2 GPU call 22: O through shared memory + TMA store in attn_tc3: parity, fuzz, timings (A/B with IEF_TC3_O_TMA=0), trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 -x > gpurun_out/r2c22_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2c22_tests.log
for t in fuzz_attn fuzz_attn_rows; do timeout 300 python tools/$t.py 13 80 2>/dev/null | tail -1; done
timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c22_bench_otma.jsonl 2>&1; echo "otma exit $?"
IEF_TC3_O_TMA=0 timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c22_bench_direct.jsonl 2>&1; echo "direct exit $?"
for v in otma direct; do echo "--- $v"; grep -h tcgen05 gpurun_out/r2c22_bench_$v.jsonl | cut -c1-60,128-190; done
V=image_editing_framework_b200/csrc/build/variants
IEF_LIB_PATH=$V/libief_b200_trace.so timeout 120 python tools/tc3_trace.py 4 20 1024 64 | tail -3; IEF_LIB_PATH=$V/libief_b200_trace.so timeout 120 python tools/tc3_trace.py 4 8 4096 40 | tail -3
