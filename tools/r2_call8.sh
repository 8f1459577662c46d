#!/usr/bin/env bash
# Round-2 GPU call 8 (2 GPUs): two-device tests, the torchrun sweep test, stored-map timings, sweep scaling 1 vs 2 GPUs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 -k "two_devices or probs or store" > gpurun_out/r2c8_kernels.log 2>&1; echo "kernels exit $?"; tail -4 gpurun_out/r2c8_kernels.log
timeout 1500 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider --timeout=900 > gpurun_out/r2c8_e2e.log 2>&1; echo "e2e exit $?"; tail -6 gpurun_out/r2c8_e2e.log
timeout 300 python tools/bench_hbm.py > gpurun_out/r2c8_hbm.jsonl 2>&1; grep -h "self-attn" gpurun_out/r2c8_hbm.jsonl | cut -c1-200
timeout 900 python tools/sweep.py --images 16 > gpurun_out/r2c8_sweep_1gpu.json 2> gpurun_out/r2c8_sweep_1gpu.err; echo "sweep1 exit $?"; cat gpurun_out/r2c8_sweep_1gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/sweep.py --images 16 > gpurun_out/r2c8_sweep_2gpu.json 2> gpurun_out/r2c8_sweep_2gpu.err; echo "sweep2 exit $?"; cat gpurun_out/r2c8_sweep_2gpu.json
