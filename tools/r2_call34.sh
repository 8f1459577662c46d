#!/usr/bin/env bash
# Round-2 GPU call 34: unclamped, range-tracked FMA-pipe exponentials in the unshifted loop
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 -x -k "tcgen05 or fp16 or row_sources or masactrl or lazy or auto or strided or workspace or full_size or key_bias or probs" > gpurun_out/r2c34_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2c34_tests.log
for t in fuzz_attn fuzz_attn_rows; do timeout 300 python tools/$t.py 23 80 2>/dev/null | tail -1; done
for v in default notrack track3; do
  if [ $v = default ]; then L=""; else L="IEF_LIB_PATH=$V/libief_b200_$v.so"; fi
  env $L timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c34_bench_$v.jsonl 2>&1; echo "$v exit $?"
done
for v in default notrack track3; do echo "--- $v"; grep -h tcgen05 gpurun_out/r2c34_bench_$v.jsonl | cut -c1-60,128-190; done
