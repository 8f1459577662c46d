#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=900 -x > gpurun_out/r02_final_tests.log 2>&1; echo "suite exit $?"; tail -3 gpurun_out/r02_final_tests.log
IEF_TC_VERSION=2 timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 -k "tcgen05 or fp16 or row_sources or masactrl or lazy or strided" > gpurun_out/r2c28_tests_v2.log 2>&1; echo "v2 tests exit $?"; tail -2 gpurun_out/r2c28_tests_v2.log
timeout 300 python tools/bench_attn.py tcgen05 nosdpa 2>/dev/null | grep tcgen05 | cut -c1-60,128-190
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
