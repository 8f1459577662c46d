#!/usr/bin/env bash
# Round-2 GPU call 47: the persistent path under the other loop variants (exact-only, fp16, one item per CTA), fuzz
mkdir -p gpurun_out
K="tcgen05 or fp16 or row_sources or masactrl or lazy or strided or key_bias"
run() { name=$1; echo "=== $name"; timeout 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --timeout=120 -k "$K" 2>&1 | tail -2; }
run default
IEF_TC3_NOMAX=0 run exact_only
IEF_TC3_PERSIST=0 run one_item_per_cta
IEF_TC_SPLITKV=0 run pair
IEF_TC3_NO_SUM_MMA=1 run no_summma
for e in "" IEF_TC3_NOMAX=0 IEF_TC_SPLITKV=0; do for t in fuzz_attn fuzz_attn_rows; do echo "fuzz $e $t: $(env $e timeout 300 python tools/$t.py 47 60 2>/dev/null | tail -1)"; done; done
