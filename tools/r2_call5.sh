#!/usr/bin/env bash
# Round-2 GPU call 5: cross-attention edit / store on the tensor pipe (cross_tc_edit.cu): parity, fuzz, timings vs the mma.sync kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 -x -k "cross" > gpurun_out/r2c5_cross_tests.log 2>&1; echo "cross tests exit $?"; tail -15 gpurun_out/r2c5_cross_tests.log
timeout 300 python tools/fuzz_cross.py 7 80 > gpurun_out/r2c5_fuzz_cross.log 2>&1; echo "fuzz exit $?"; tail -3 gpurun_out/r2c5_fuzz_cross.log
timeout 300 python tools/bench_hbm.py > gpurun_out/r2c5_hbm_tc_edit.jsonl 2>&1; echo "hbm exit $?"
IEF_CROSS_TC_EDIT=0 timeout 300 python tools/bench_hbm.py > gpurun_out/r2c5_hbm_mma_edit.jsonl 2>&1; echo "hbm mma exit $?"
grep -h cross_attn gpurun_out/r2c5_hbm_tc_edit.jsonl | cut -c1-170; echo ---; grep -h cross_attn gpurun_out/r2c5_hbm_mma_edit.jsonl | cut -c1-170
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider --timeout=600 > gpurun_out/r2c5_e2e.log 2>&1; echo "e2e exit $?"; tail -8 gpurun_out/r2c5_e2e.log
