#!/usr/bin/env bash
# SASS opcode evidence per object file of libief_b200.so (runs on the CPU box: cuobjdump only). Usage: tools/sass_histogram.sh > profiles/rNN_sass_histogram.txt
# UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA load, HMMA = mma.sync (legacy tensor path), LDGSTS = cp.async.
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
b="$here/image_editing_framework_b200/csrc/build"
printf "%-18s %8s %8s %8s %8s %8s %8s %8s %8s\n" object UTCHMMA LDTM STTM UTMALDG UTMASTG HMMA LDGSTS MUFU.EX2
for f in attn_tc3 attn_tc2 attn_tc cross_tc cross_tc_edit attn_mma cross_attn cross_attn_bwd elementwise; do
  s="$(cuobjdump -sass "$b/$f.o")"
  c() { grep -c "$1" <<< "$s"; }
  printf "%-18s %8s %8s %8s %8s %8s %8s %8s %8s\n" "$f.cu" "$(c 'UTCHMMA')" "$(c 'LDTM')" "$(c 'STTM')" "$(c 'UTMALDG')" "$(c 'UTMASTG')" "$(c ' HMMA')" "$(c 'LDGSTS')" "$(c 'MUFU.EX2')"
done
echo
echo "# whole library"
s="$(cuobjdump -sass "$here/image_editing_framework_b200/libief_b200.so")"
for op in UTCHMMA LDTM STTM UTMALDG HMMA LDGSTS; do printf "%-10s %s\n" $op "$(grep -c "$op" <<< "$s")"; done
