#!/usr/bin/env bash
# Round-2 GPU call 48: the remainder of a hybrid call as a programmatic dependent launch (IEF_TC3_PDL)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=120 -x -k "tcgen05 or fp16 or row_sources or masactrl or lazy or strided or key_bias or full_size" 2>&1 | tail -2
IEF_TC_SPLITKV=2 timeout 300 python tools/fuzz_attn.py 48 60 2>/dev/null | tail -1
timeout 300 python tools/fuzz_attn_rows.py 48 60 2>/dev/null | tail -1
for e in IEF_TC3_PDL=1 IEF_TC3_PDL=0; do echo "--- $e"; env $e timeout 300 python tools/bench_attn.py tcgen05 big nosdpa 2>/dev/null | grep tcgen05 | cut -c11-20,100-160; done
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider --timeout=600 -x 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c48_bench.json 2> gpurun_out/r2c48_bench.err; echo "bench exit $?"; cut -c1-200 gpurun_out/r2c48_bench.json; tail -2 gpurun_out/r2c48_bench.err
