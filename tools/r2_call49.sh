#!/usr/bin/env bash
# Round-2 GPU call 49: launch modes with the dependent-launch remainder: auto vs forced pair vs forced hybrid (same box), SDPA beside
for e in IEF_TC_SPLITKV=-1 IEF_TC_SPLITKV=0 IEF_TC_SPLITKV=2; do echo "--- $e"; env $e timeout 300 python tools/bench_attn.py tcgen05 big nosdpa 2>/dev/null | grep tcgen05 | cut -c11-20,100-160; done
timeout 300 python tools/bench_attn.py tcgen05 big 2>/dev/null | grep sdpa | cut -c1-110
