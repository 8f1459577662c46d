#!/usr/bin/env bash
# Round-2 GPU call 3: the rewritten bench (public API, GPU reference leg, reference arm) + overflow diagnostic + the inversion replay test.
mkdir -p gpurun_out
timeout 200 python tools/diag/overflow_case.py > gpurun_out/r2c3_diag.txt 2>&1; echo "diag exit $?"
IEF_TC3_NOMAX=0 timeout 200 python tools/diag/overflow_case.py >> gpurun_out/r2c3_diag.txt 2>&1; echo "diag exit $?"
cut -c1-420 gpurun_out/r2c3_diag.txt
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider --timeout=300 -k "inversion" > gpurun_out/r2c3_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2c3_tests.log
timeout 900 python bench.py > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err; echo "bench exit $?"; tail -5 gpurun_out/r2c3_bench.err; cat gpurun_out/r2c3_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c3_bench_ref.json 2> gpurun_out/r2c3_bench_ref.err; echo "ref exit $?"; cat gpurun_out/r2c3_bench_ref.json
