#!/usr/bin/env bash
# Round-2 GPU call 39: cost of the two parts of a hybrid launch (full waves persistent | remainder split-KV)
for part in 0 1 2; do echo "--- IEF_TC3_DIAG_PART=$part (hybrid forced)"; IEF_TC_SPLITKV=2 IEF_TC3_DIAG_PART=$part timeout 300 python tools/bench_attn.py tcgen05 big nosdpa 2>&1 | grep tcgen05 | cut -c1-60,100-190; done
