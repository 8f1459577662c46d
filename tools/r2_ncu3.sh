#!/usr/bin/env bash
# Round-2 ncu evidence for the persistent attn_tc3 (ONE gpurun call, one GPU). Every ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
full() {  # name, kernel regex, count, command...
  name=$1; rx=$2; cnt=$3; shift 3
  "$@" > gpurun_out/r02_plain_$name.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c $cnt -o gpurun_out/r02_prof_$name -f "$@" > gpurun_out/r02_ncu_$name.log 2>&1
  echo "$name exit $?"
}
full attn_tc3_persist_sd15 attn_tc3 2 python tools/profile_attn.py 4 8 4096 40 tc
full attn_tc3_persist_d64 attn_tc3 2 python tools/profile_attn.py 4 10 4096 64 tc
BENCH="python bench.py --steps 1 --warmup 0 --ddim-steps 3 --no-cpu-baseline --no-graphs"
$BENCH > gpurun_out/r02_bench_short_plain2.json 2> gpurun_out/r02_bench_short_plain2.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 5000 -c 3000 --csv --log-file gpurun_out/r02_bench_launch_list_persistent.csv $BENCH > gpurun_out/r02_ncu_launches2.log 2>&1
echo "launch list exit $?"
ls -la gpurun_out/*persist*.ncu-rep gpurun_out/r02_bench_launch_list_persistent.csv
