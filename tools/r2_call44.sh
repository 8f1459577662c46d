#!/usr/bin/env bash
# Round-2 GPU call 44: full GPU suite, kernel microbench with library SDPA beside it, smoke, bench.py at the persistent kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=900 -x > gpurun_out/r02_persist_tests.log 2>&1; echo "suite exit $?"; tail -3 gpurun_out/r02_persist_tests.log
timeout 300 python tools/bench_attn.py tcgen05 > gpurun_out/r02_kernel_microbench_persistent.jsonl 2>/dev/null; grep -h "tcgen05\|sdpa" gpurun_out/r02_kernel_microbench_persistent.jsonl | cut -c1-60,100-190
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_persist_bench.json 2> gpurun_out/r02_persist_bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/r02_persist_bench.json
