#!/usr/bin/env bash
# ncu evidence for profiles/: per-launch device time of ONE MasaCtrl-controlled B=4 UNet forward + step update inside the bench command
# (eager mode so the NVTX range filter applies; cold-cache / serialised replay: compare SHARES, not absolute times).
# The ncu run follows a plain run of the same command that exited 0. ~0.3 s per profiled launch: keep -c small.
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 0 --ddim-steps 5 --no-cpu-baseline --no-graphs"
$BENCH > gpurun_out/bench_plain_short.json 2> gpurun_out/bench_plain_short.err && \
timeout 1500 ncu --nvtx --nvtx-include "edit_ctrl/" --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv \
  --log-file gpurun_out/launches_bench.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
