#!/usr/bin/env bash
# Round-2 GPU call 15: FMA-pipe exponentials inside the exp section, now that the MMA issue chain is out of the way
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
for v in emul6 emul4 emul3 emul2; do
  IEF_LIB_PATH=$V/libief_b200_$v.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c15_bench_$v.jsonl 2>&1; echo "$v exit $?"
  IEF_LIB_PATH=$V/libief_b200_$v.so timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 -k "tcgen05 or fp16 or row_sources or masactrl or full_size" > gpurun_out/r2c15_tests_$v.log 2>&1; echo "$v tests exit $?"; tail -4 gpurun_out/r2c15_tests_$v.log | cut -c1-200
done
for v in emul6 emul4 emul3 emul2; do echo "--- $v"; grep -h tcgen05 gpurun_out/r2c15_bench_$v.jsonl | cut -c1-190; done
