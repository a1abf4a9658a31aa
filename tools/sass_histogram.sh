#!/bin/bash
# SASS evidence for the shipped library: which kernels contain tcgen05 (UTC*MMA), TMEM loads (LDTM), TMA (UTMA*) and
# mbarrier (SYNCS) instructions.  usage: tools/sass_histogram.sh > profiles/sass_histogram_rNN.txt
SO=${1:-structure_knowledge_distillation_b200/libskd_b200.so}
TMP=$(mktemp)
cuobjdump -sass "$SO" > "$TMP"
echo "# cuobjdump -sass $SO : opcode histogram (whole library)"
grep -oE 'UTC[A-Z0-9.]*MMA[A-Z0-9.]*|LDTM[A-Z0-9.]*|STTM[A-Z0-9.]*|UTMA[A-Z0-9.]*|UTCBAR[A-Z0-9.]*|UTCCP[A-Z0-9.]*|SYNCS[A-Z0-9.]*|UBLKCP[A-Z0-9.]*|HMMA[A-Z0-9.]*|LDGMC[A-Z0-9.]*|STG[A-Z0-9.]*MC[A-Z0-9.]*|REDG?MC[A-Z0-9.]*' "$TMP" | sort | uniq -c | sort -rn
echo
echo "# per kernel: UTCHMMA / LDTM / UTMALDG / UTMASTG counts (kernels with at least one)"
awk '/Function :/ {name=$3} /UTCHMMA/ {m[name]++} /LDTM/ {l[name]++} /UTMALDG/ {t[name]++} /UTMASTG/ {s[name]++}
     END {for (k in m) printf "%5d %5d %5d %5d  %s\n", m[k], l[k], t[k], s[k], k}' "$TMP" | sort -rn | while read a b c d n; do
  printf "%5s %5s %5s %5s  %s\n" "$a" "$b" "$c" "$d" "$(echo "$n" | c++filt | cut -c1-150)"; done
rm -f "$TMP"
