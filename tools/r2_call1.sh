#!/usr/bin/env bash
# Round-2 GPU call 1: parity of the unshifted-softmax generation of attn_tc3, then A/B timings of its compile-time variants.
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 -x -k "tcgen05 or fp16 or row_sources or masactrl or lazy or auto or strided or workspace or full_size or key_bias" > gpurun_out/r2c1_tests.log 2>&1
echo "tests exit $?"; tail -5 gpurun_out/r2c1_tests.log
timeout 300 python tools/bench_attn.py tcgen05 > gpurun_out/r2c1_bench_default.jsonl 2> gpurun_out/r2c1_bench_default.err; echo "default exit $?"
IEF_TC3_NOMAX=0 timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c1_bench_exact.jsonl 2>&1; echo "exact exit $?"
for v in freerun scaleturn late freescale; do
  IEF_LIB_PATH=$V/libief_b200_$v.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c1_bench_$v.jsonl 2>&1; echo "$v exit $?"
done
IEF_LIB_PATH=$V/libief_b200_trace.so timeout 120 python tools/tc3_trace.py 4 8 4096 40 > gpurun_out/r2c1_trace_d40.txt 2>&1; echo "trace40 exit $?"
IEF_LIB_PATH=$V/libief_b200_trace.so timeout 120 python tools/tc3_trace.py 4 10 4096 64 > gpurun_out/r2c1_trace_d64.txt 2>&1; echo "trace64 exit $?"
grep -h tcgen05 gpurun_out/r2c1_bench_*.jsonl | cut -c1-200
