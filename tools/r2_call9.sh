#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider --timeout=600 -k "masked_variants or custom_reference" > gpurun_out/r2c9_e2e.log 2>&1; echo "e2e exit $?"; tail -4 gpurun_out/r2c9_e2e.log
bash tools/r2_ncu.sh
