#!/usr/bin/env bash
# ncu evidence for profiles/: (1) launch list of the bench command (per-launch device time, cold-cache/serialised: compare SHARES),
# (2) one --set full capture of the dominant kernel. Each ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 0 --ddim-steps 2 --no-cpu-baseline"
$BENCH > gpurun_out/bench_plain_short.json 2> gpurun_out/bench_plain_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 2300 --csv --log-file gpurun_out/launches_bench.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
python tools/profile_attn.py 4 8 4096 40 tc > gpurun_out/pa_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_tc2_kernel -s 4 -c 2 -o gpurun_out/prof_tc2 python tools/profile_attn.py 4 8 4096 40 tc > gpurun_out/ncu_tc2.log 2>&1
echo "set full exit $?"
