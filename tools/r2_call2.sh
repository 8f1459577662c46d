#!/usr/bin/env bash
# Round-2 GPU call 2: in-turn FMA-pipe exponentials (A/B) + the overflow diagnostic.
mkdir -p gpurun_out
V=image_editing_framework_b200/csrc/build/variants
timeout 200 python tools/diag/overflow_case.py > gpurun_out/r2c2_diag.txt 2>&1; echo "diag exit $?"
IEF_TC3_NOMAX=0 timeout 200 python tools/diag/overflow_case.py >> gpurun_out/r2c2_diag.txt 2>&1; echo "diag exit $?"
cat gpurun_out/r2c2_diag.txt | cut -c1-400
for v in emul2 emul3 emul4 emul6; do
  IEF_LIB_PATH=$V/libief_b200_$v.so timeout 300 python tools/bench_attn.py tcgen05 big nosdpa > gpurun_out/r2c2_bench_$v.jsonl 2>&1; echo "$v exit $?"
  IEF_LIB_PATH=$V/libief_b200_$v.so timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 -k "tcgen05 or fp16 or row_sources or masactrl or full_size" > gpurun_out/r2c2_tests_$v.log 2>&1; echo "$v tests exit $?"; tail -2 gpurun_out/r2c2_tests_$v.log
done
grep -h tcgen05 gpurun_out/r2c2_bench_*.jsonl | cut -c1-200
