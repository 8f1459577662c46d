#!/usr/bin/env bash
# Round-2 ncu evidence (ONE gpurun call, one GPU): launch list of a short eager bench run + `--set full` captures of the kernels VERDICT asks for.
# Every ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 0 --ddim-steps 3 --no-cpu-baseline --no-graphs"
$BENCH > gpurun_out/r02_bench_short_plain.json 2> gpurun_out/r02_bench_short_plain.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_bench_launch_list.csv $BENCH > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list exit $?"
full() {  # name, kernel regex, count, command...
  name=$1; rx=$2; cnt=$3; shift 3
  "$@" > gpurun_out/r02_plain_$name.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c $cnt -o gpurun_out/r02_prof_$name -f "$@" > gpurun_out/r02_ncu_$name.log 2>&1
  echo "$name exit $?"
}
full attn_tc3_sd15 attn_tc3 2 python tools/profile_attn.py 4 8 4096 40 tc
full attn_tc3_d64 attn_tc3 2 python tools/profile_attn.py 4 10 4096 64 tc
full cross_4096_40 cross_tc 3 python tools/profile_cross.py 4096 40
full attn_mma_256_160 attn_mma 1 python tools/profile_attn.py 4 8 256 160 mma
full probs_1024_80 probs_from_lse 1 python tools/profile_probs.py
ls -la gpurun_out/*.ncu-rep
