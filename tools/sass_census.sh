#!/bin/bash
# SASS census of libcleanba_b200.so: per kernel, the Blackwell tensor-core / TMEM / TMA opcodes that prove the path
# (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk 1-D TMA, UTMALDG = tensor-map TMA, SYNCS = mbarrier ops).
# Usage: tools/sass_census.sh > profiles/rNN_sass_census.txt        (runs on the build box: cuobjdump needs no GPU)
SO=${1:-cleanba_b200/libcleanba_b200.so}
echo "# cuobjdump -sass $SO  ($(date -u +%Y-%m-%dT%H:%MZ), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
printf "%-72s %8s %6s %7s %8s %6s %6s\n" kernel UTCHMMA LDTM UBLKCP UTMALDG SYNCS HMMA
cuobjdump -sass "$SO" | awk '
  /Function :/ { if (name != "") printf "%-72s %8d %6d %7d %8d %6d %6d\n", name, m, l, b, t, s, h; name=$3; m=l=b=t=s=h=0 }
  /UTCHMMA/ {m++} /LDTM/ {l++} /UBLKCP/ {b++} /UTMALDG/ {t++} /SYNCS/ {s++} / HMMA/ {h++}
  END { printf "%-72s %8d %6d %7d %8d %6d %6d\n", name, m, l, b, t, s, h }' | while read -r n rest; do printf "%-72s %s\n" "$(echo $n | c++filt | cut -c1-72)" "$rest"; done
echo
echo "# excerpt: the MMA issue loop of the 16->16 channel conv (k_conv_umma<2,16>)"
cuobjdump -sass "$SO" | awk '/Function : .*k_conv_ummaILi2ELi16E/ {on=1} on && /UTCHMMA|UBLKCP|LDTM|SYNCS|UTCBAR/ {print} /Function :/ && !/k_conv_ummaILi2ELi16E/ {on=0}' | head -60
