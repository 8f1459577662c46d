#!/usr/bin/env bash
# Round-2 GPU call 52: HEAD verification — full GPU suite, smoke, bench.py
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout=900 -x > gpurun_out/r02_head_tests.log 2>&1; echo "suite exit $?"; tail -2 gpurun_out/r02_head_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_head_bench.json 2> gpurun_out/r02_head_bench.err; echo "bench exit $?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_head_bench.json').read().strip().splitlines()[-1])
print(l['value'], l['e2e']['value'], l['roofline']['achieved'], l['roofline']['frac'], l['gpu_launches'], l['clocks'])
PY
