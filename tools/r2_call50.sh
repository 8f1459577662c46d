#!/usr/bin/env bash
# Round-2 GPU call 50: final state — full GPU suite, kernel microbench with library SDPA beside it, smoke, bench.py
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=900 -x > gpurun_out/r02_final2_tests.log 2>&1; echo "suite exit $?"; tail -3 gpurun_out/r02_final2_tests.log
timeout 300 python tools/bench_attn.py tcgen05 > gpurun_out/r02_kernel_microbench_final.jsonl 2>/dev/null; grep -h "tcgen05\|sdpa" gpurun_out/r02_kernel_microbench_final.jsonl | cut -c1-30,90-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_final2_bench.json 2> gpurun_out/r02_final2_bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/r02_final2_bench.json
