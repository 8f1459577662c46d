#!/usr/bin/env bash
# Round-2 GPU call 4: full-geometry parity goldens, consistent row sums (A/B), kernel tests.
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider --timeout=600 -k "baseline_attention_geometry" -s > gpurun_out/r2c4_fullgeo.log 2>&1; echo "fullgeo exit $?"; grep -E "worst layer|passed|failed|Error|assert" gpurun_out/r2c4_fullgeo.log | head -40
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 > gpurun_out/r2c4_kernels.log 2>&1; echo "kernels exit $?"; tail -5 gpurun_out/r2c4_kernels.log
timeout 200 python tools/diag/overflow_case.py > gpurun_out/r2c4_diag.txt 2>&1; cut -c1-200 gpurun_out/r2c4_diag.txt
timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c4_bench_default.jsonl 2>&1; echo "default exit $?"
IEF_LIB_PATH=$V/libief_b200_rnsum.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c4_bench_rnsum.jsonl 2>&1; echo "rnsum exit $?"
grep -h tcgen05 gpurun_out/r2c4_bench_*.jsonl | cut -c1-200
