#!/usr/bin/env bash
# Round-2 GPU call 6: whole GPU suite after the LocalBlend / sweep / golden changes + a short full-size sweep.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider --timeout=900 -s > gpurun_out/r2c6_e2e.log 2>&1; echo "e2e exit $?"; grep -E "worst layer|passed|failed|Error|assert " gpurun_out/r2c6_e2e.log | head -40
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fuzz.py -q -m gpu -p no:cacheprovider --timeout=600 > gpurun_out/r2c6_kernels.log 2>&1; echo "kernels exit $?"; tail -4 gpurun_out/r2c6_kernels.log
timeout 900 python tools/sweep.py --images 12 --out gpurun_out/r2c6_sweep_records.jsonl > gpurun_out/r2c6_sweep.json 2> gpurun_out/r2c6_sweep.err; echo "sweep exit $?"; tail -3 gpurun_out/r2c6_sweep.err; cat gpurun_out/r2c6_sweep.json
