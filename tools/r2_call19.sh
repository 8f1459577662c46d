#!/usr/bin/env bash
# Round-2 GPU call 19: final single-GPU numbers of generation 3d: kernel microbench with SDPA as context, both bench arms, HBM kernels, one ncu capture
mkdir -p gpurun_out
timeout 400 python tools/bench_attn.py > gpurun_out/r02_kernel_microbench_gen3d.jsonl 2> gpurun_out/r02_kernel_microbench_gen3d.err; echo "microbench exit $?"
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err; echo "ref exit $?"
timeout 300 python tools/bench_hbm.py > gpurun_out/r02_hbm_kernels.jsonl 2>&1; echo "hbm exit $?"
timeout 300 python tools/bench_cross.py > gpurun_out/r02_cross_plain.jsonl 2>&1; echo "cross exit $?"
python tools/profile_attn.py 4 8 4096 40 tc > gpurun_out/r02_plain_3d.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc3 -s 4 -c 2 -o gpurun_out/r02_prof_attn_tc3_3d_sd15 -f python tools/profile_attn.py 4 8 4096 40 tc > gpurun_out/r02_ncu_3d.log 2>&1; echo "ncu exit $?"
grep -h tcgen05 gpurun_out/r02_kernel_microbench_gen3d.jsonl | cut -c1-60,128-190; python -c "
import json; d=json.load(open('gpurun_out/r02_final_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['gpu_reference_baseline']['value'], d['cpu_baseline']['value'])"
