#!/usr/bin/env bash
# Runs the GPU test groups in separate processes (a faulting kernel poisons only its own process).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
run() { name=$1; shift; echo "=== $name" ; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=300 -k "$1" > gpurun_out/t_$name.log 2>&1; echo "exit $?"; tail -5 gpurun_out/t_$name.log; }
run probe "umma_probe"
run mma "attn_mma or probs_out or rejects"
run tc "tcgen05 or fp16 or row_sources or masactrl or lazy or auto or strided"
IEF_TC_VERSION=2 run tc_v2 "tcgen05 or fp16 or row_sources or masactrl or lazy or strided"
IEF_TC_SPLITKV=1 run tc_v3_split "tcgen05 or fp16 or row_sources or masactrl or lazy or strided"
IEF_TC_SPLITKV=2 run tc_v3_hybrid "tcgen05 or fp16 or row_sources or masactrl or lazy or strided"
IEF_TC_SPLITKV=0 run tc_v3_pair "tcgen05 or fp16 or row_sources or masactrl or lazy or strided"
IEF_TC3_PERSIST=0 run tc_v3_one_item_per_cta "tcgen05 or fp16 or row_sources or masactrl or lazy or strided"
IEF_TC_VERSION=1 run tc_v1 "tcgen05 or fp16 or row_sources or masactrl or lazy or strided"
IEF_TC3_NO_SUM_MMA=1 run tc_v3_no_summma "tcgen05 or fp16 or row_sources or masactrl or lazy or strided"
IEF_TC3_SKIPMAX=2 run tc_v3_skip_everywhere "tcgen05 or fp16 or row_sources or masactrl or lazy or strided"
IEF_TC3_NOMAX=0 run tc_v3_exact_only "tcgen05 or fp16 or row_sources or masactrl or lazy or strided"
run cross "cross_attention"
IEF_CROSS_TC=0 run cross_mma_only "cross_attention"
IEF_CROSS_TC_EDIT=0 run cross_edit_on_mma "cross_attention"
IEF_CROSS_TC_ONE_LAUNCH=0 run cross_edit_two_launches "cross_attention"
IEF_CROSS_TC_ONE_LAUNCH=1 run cross_edit_one_launch "cross_attention"
IEF_CROSS_TC_ONE_LAUNCH=2 run cross_edit_dependent_launch "cross_attention"
run masked "key_bias or mask_blend"
IEF_PROBS_VIA_LSE=0 run probs_two_sweep "probs_out"
run backward "cross_attention_backward"
run elem "ddim or accumulate or local_blend"
echo "=== e2e"; timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -p no:cacheprovider --timeout=600 > gpurun_out/t_e2e.log 2>&1; echo "exit $?"; tail -25 gpurun_out/t_e2e.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -3 gpurun_out/smoke.log
fuzz() { name=$1; shift; echo "=== fuzz $name"; for t in fuzz_attn fuzz_attn_rows fuzz_cross; do timeout 300 python tools/$t.py 7 60 2> gpurun_out/f_${name}_$t.err | tail -2; echo "exit ${PIPESTATUS[0]}"; done; }
fuzz default
IEF_TC_VERSION=2 fuzz v2
IEF_TC_SPLITKV=0 fuzz v3_pair
IEF_TC_SPLITKV=1 fuzz v3_split
IEF_TC3_SKIPMAX=2 fuzz skip_everywhere
IEF_TC3_NOMAX=0 fuzz exact_only
IEF_TC3_PERSIST=0 fuzz one_item_per_cta
IEF_PROBS_VIA_LSE=0 fuzz probs_two_sweep
IEF_CROSS_TC_EDIT=0 fuzz cross_edit_on_mma
IEF_CROSS_TC_ONE_LAUNCH=0 fuzz cross_edit_two_launches
IEF_TC3_NO_SUM_MMA=1 fuzz no_summma
