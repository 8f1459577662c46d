#!/usr/bin/env bash
# Round-2 GPU call 14: two MMA issuer warps (one per stream): parity, timings vs the single issuer, no-exp diagnostic, trace
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 -x -k "tcgen05 or fp16 or row_sources or masactrl or lazy or auto or strided or workspace or full_size or key_bias or probs" > gpurun_out/r2c14_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/r2c14_tests.log
for t in fuzz_attn fuzz_attn_rows; do timeout 300 python tools/$t.py 11 60 2>/dev/null | tail -1; done
timeout 300 python tools/bench_attn.py tcgen05 > gpurun_out/r2c14_bench_dual.jsonl 2> gpurun_out/r2c14_bench_dual.err; echo "dual exit $?"
IEF_LIB_PATH=$V/libief_b200_single.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c14_bench_single.jsonl 2>&1; echo "single exit $?"
IEF_LIB_PATH=$V/libief_b200_skip1.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c14_bench_skip1.jsonl 2>&1; echo "skip1 exit $?"
IEF_LIB_PATH=$V/libief_b200_trace.so timeout 120 python tools/tc3_trace.py 4 8 4096 40 > gpurun_out/r2c14_trace_d40.txt 2>&1
IEF_LIB_PATH=$V/libief_b200_trace.so timeout 120 python tools/tc3_trace.py 4 10 4096 64 > gpurun_out/r2c14_trace_d64.txt 2>&1
for f in dual single skip1; do echo "--- $f"; grep -h '"impl": "tcgen05"\|sdpa' gpurun_out/r2c14_bench_$f.jsonl | cut -c1-190; done
sed -n '1,1p;6,10p' gpurun_out/r2c14_trace_d40.txt; tail -2 gpurun_out/r2c14_trace_d40.txt; sed -n '6,10p' gpurun_out/r2c14_trace_d64.txt; tail -2 gpurun_out/r2c14_trace_d64.txt
