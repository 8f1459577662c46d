#!/usr/bin/env bash
# Round-2 GPU call 16: whole GPU suite + smoke + bench with the new attn_tc3 default (two issuers, 1/4 of the exponentials on the FMA pipe)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=900 -x > gpurun_out/r2c16_gpu_suite.log 2>&1; echo "suite exit $?"; tail -4 gpurun_out/r2c16_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c16_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2c16_smoke.log
timeout 900 python bench.py > gpurun_out/r2c16_bench.json 2> gpurun_out/r2c16_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2c16_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['gpu_reference_baseline']['value'], d['cpu_baseline']['value'], d['clocks'])"
IEF_LIB_PATH=image_editing_framework_b200/csrc/build/variants/libief_b200_trace.so timeout 120 python tools/tc3_trace.py 4 8 4096 40 > gpurun_out/r2c16_trace_d40.txt 2>&1; sed -n '1,1p;6,10p' gpurun_out/r2c16_trace_d40.txt; tail -2 gpurun_out/r2c16_trace_d40.txt
