#!/usr/bin/env bash
# Round-2 GPU call 25: whole GPU suite, smoke and both bench arms at HEAD (attn_tc3: TMA-store epilogue)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=900 -x > gpurun_out/r02_final_tests.log 2>&1; echo "suite exit $?"; tail -3 gpurun_out/r02_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c25_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2c25_smoke.log
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02_final_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['gpu_reference_baseline']['value'], d['cpu_baseline']['value'], d['gpu_launches'])"
timeout 400 python tools/bench_attn.py > gpurun_out/r02_kernel_microbench_gen3d.jsonl 2>/dev/null; grep -h "tcgen05\|sdpa" gpurun_out/r02_kernel_microbench_gen3d.jsonl | cut -c1-60,100-190
