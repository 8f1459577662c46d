#!/usr/bin/env bash
# Round-2 multi-GPU numbers: bash tools/r2_scale.sh N  -> bench.py (both arms at N) and the 48-image sweep on N GPUs of one box
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671"; fi
timeout 900 $L bench.py --gpus $N > gpurun_out/r02_scale_bench_${N}gpu.json 2> gpurun_out/r02_scale_bench_${N}gpu.err; echo "bench exit $?"; cut -c1-260 gpurun_out/r02_scale_bench_${N}gpu.json
timeout 1500 $L tools/sweep.py --images 48 --out gpurun_out/r02_scale_sweep_${N}gpu_records.jsonl > gpurun_out/r02_scale_sweep_${N}gpu.json 2> gpurun_out/r02_scale_sweep_${N}gpu.err; echo "sweep exit $?"; grep "^{" gpurun_out/r02_scale_sweep_${N}gpu.json | cut -c1-600
