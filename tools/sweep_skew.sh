#!/usr/bin/env bash
# sweep the tile-B start skew for the tcgen05 kernels (each value needs its own process: the env var is read once)
for v in 3 2; do for s in 0 600 1000 1400 1800 2400; do
  echo "version=$v skew=$s $(IEF_TC_VERSION=$v IEF_TC_SKEW=$s python tools/bench_attn.py tcgen05 2>/dev/null | grep -E 'sd15_64|big_d40|big_d64|sdxl_64' | grep -v sdpa | python -c "
import sys, json
print(' '.join(f\"{json.loads(l)['shape']}={json.loads(l)['tflops_median']}\" for l in sys.stdin))")"
done; done
