#!/usr/bin/env bash
# Round-2 GPU call 7: e2e suite after the test changes; PnP anomaly diagnostic; methods bench on sd21.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider --timeout=900 > gpurun_out/r2c7_e2e.log 2>&1; echo "e2e exit $?"; tail -6 gpurun_out/r2c7_e2e.log
timeout 300 python tools/diag/pnp_anomaly.py > gpurun_out/r2c7_pnp_anomaly.jsonl 2>&1; echo "diag exit $?"; cat gpurun_out/r2c7_pnp_anomaly.jsonl
timeout 900 python tools/bench_methods.py 50 sd21 > gpurun_out/r2c7_methods_sd21.log 2>&1; echo "sd21 exit $?"; tail -12 gpurun_out/r2c7_methods_sd21.log | cut -c1-300
