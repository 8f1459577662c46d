#!/usr/bin/env bash
# Round-2 GPU call 55: default cross-attention launch policy (dependent plain launch) — kernel tests, e2e suite (P2P under graph replay), fuzz
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=120 -x -k "cross_attention" 2>&1 | tail -1
timeout 200 python tools/fuzz_cross.py 55 40 2>/dev/null | tail -1
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider --timeout=500 -x 2>&1 | tail -1
