#!/usr/bin/env bash
# Round-2 ncu evidence, second call: the single-launch P2P cross-attention edit and the stored-map sweep at HEAD
set -u
mkdir -p gpurun_out
full() { name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  "$@" > gpurun_out/r02_plain_$name.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o gpurun_out/r02_prof_$name -f "$@" > gpurun_out/r02_ncu_$name.log 2>&1
  echo "$name exit $?"; }
full cross_edit_4096_40 cross_tc 4 2 python tools/profile_cross.py 4096 40
full cross_edit_1024_80 cross_tc 4 2 python tools/profile_cross.py 1024 80
full probs_1024_80 "probs_from_lse|attn_tc2" 2 2 python tools/profile_probs.py
ls -la gpurun_out/r02_prof_cross_edit* gpurun_out/r02_prof_probs*
