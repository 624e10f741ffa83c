#!/usr/bin/env bash
# Counts, per kernel of go-rio_b200/libapdgicp.so, the SASS mnemonics that show which memory machinery the kernel uses
# (B200_PROFILING.md: UBLKCP = cp.async.bulk (TMA 1-D), LDGSTS = cp.async, SYNCS = mbarrier, UCGABAR = cluster barrier,
# LDG.E.CONSTANT = read-only path, CCTL = cache control on acquire, ATOMS/ATOMG/RED = atomics, DFMA/DMUL/DADD = fp64).
# No GPU needed. usage: profiles/sass_counts.sh > profiles/r02_sass_counts.txt
SO="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)/go-rio_b200/libapdgicp.so"
echo "# cuobjdump -sass $(basename "$SO") ($(cuobjdump -lelf "$SO" | grep -c sm_100a) sm_100a cubin(s), no other arch)"
cuobjdump -sass "$SO" | awk '
  /Function :/ { name=$3; next }
  name != "" {
    n[name]++
    if ($0 ~ /UBLKCP/) a[name]++
    if ($0 ~ /LDGSTS/) b[name]++
    if ($0 ~ /SYNCS/) c[name]++
    if ($0 ~ /UCGABAR|CGABAR/) d[name]++
    if ($0 ~ /LDG\.E[^ ]*\.CONSTANT/) e[name]++
    if ($0 ~ /CCTL/) f[name]++
    if ($0 ~ /ATOMS|ATOMG|RED\./) g[name]++
    if ($0 ~ /DFMA|DMUL|DADD/) h[name]++
    if ($0 ~ /SHFL/) s[name]++
    if ($0 ~ /STL|LDL/) l[name]++
  }
  END {
    printf "%-70s %7s %6s %6s %5s %7s %8s %4s %5s %5s %5s %7s\n", "kernel", "instr", "UBLKCP", "LDGSTS", "SYNCS", "CGABAR", "LDG.CONST", "CCTL", "ATOM", "FP64", "SHFL", "LDL/STL"
    for (k in n) printf "%-70s %7d %6d %6d %5d %7d %8d %4d %5d %5d %5d %7d\n", substr(k,1,70), n[k], a[k], b[k], c[k], d[k], e[k], f[k], g[k], h[k], s[k], l[k]
  }' | (read -r hdr; echo "$hdr"; sort)
