#!/usr/bin/env bash
# Round-2 GPU call 27: generation-2 kernel (head_dim 65..128, and IEF_TC_VERSION=2) with one MMA issuer warp per query tile + new dispatch assertions
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 -x > gpurun_out/r2c27_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2c27_tests.log
IEF_TC_VERSION=2 timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 -x -k "tcgen05 or fp16 or row_sources or masactrl or lazy or strided" > gpurun_out/r2c27_tests_v2.log 2>&1; echo "v2 tests exit $?"; tail -2 gpurun_out/r2c27_tests_v2.log
for t in fuzz_attn fuzz_attn_rows; do timeout 300 python tools/$t.py 17 60 2>/dev/null | tail -1; IEF_TC_VERSION=2 timeout 300 python tools/$t.py 19 60 2>/dev/null | tail -1; done
timeout 300 python tools/bench_attn.py tcgen05 nosdpa 2>/dev/null | grep tcgen05 | cut -c1-60,128-190
IEF_TC_VERSION=2 timeout 300 python tools/bench_attn.py tcgen05 big nosdpa 2>/dev/null | grep tcgen05 | cut -c1-60,128-190
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider --timeout=900 -x -k "baseline_attention_geometry" > gpurun_out/r2c27_e2e.log 2>&1; echo "e2e exit $?"; tail -2 gpurun_out/r2c27_e2e.log
