#!/usr/bin/env bash
# Round-2 GPU call 54: P2P cross-attention edit: one launch | two serialised launches | edit launch + dependent plain launch
for m in 1 0 2; do echo "--- IEF_CROSS_TC_ONE_LAUNCH=$m"; IEF_CROSS_TC_ONE_LAUNCH=$m timeout 200 python tools/bench_hbm.py 2>/dev/null | grep "cross_attn" | cut -c1-140; done
IEF_CROSS_TC_ONE_LAUNCH=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=120 -x -k "cross_attention" 2>&1 | tail -2
IEF_CROSS_TC_ONE_LAUNCH=2 timeout 200 python tools/fuzz_cross.py 54 40 2>/dev/null | tail -1
