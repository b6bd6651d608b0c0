#!/bin/bash
# Per-kernel counts of the SASS opcodes that prove the Blackwell-native paths (tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM,
# TMA = UBLKCP / UTMALDG / UTMASTG / UTMAPF, cp.async = LDGSTS, 3-input min = FMNMX3) in the built library.
#   tools/sass_opcodes.sh > profiles/sass_opcodes.txt
cd "$(dirname "$0")/.."
LIB=3d_multiview_reg_b200/liblmpcr_b200.so
echo "# cuobjdump -sass $LIB | per-kernel opcode counts ($(date -u +%F), $(git rev-parse --short HEAD))"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3; next }
  { for (i=1;i<=NF;i++) { op=$i; sub(/\..*/,"",op);
      if (op=="UTCHMMA"||op=="LDTM"||op=="STTM"||op=="UBLKCP"||op=="UTMALDG"||op=="UTMASTG"||op=="UTMAPF"||op=="LDGSTS"||op=="FMNMX3"||op=="UTCBAR"||op=="ELECT") c[fn" "op]++ } }
  END { for (k in c) print k, c[k] }' | sort | c++filt 2>/dev/null | sed 's/(anonymous namespace):://g; s/lmpcr:://g; s/^void //' | awk '{ n=$NF; op=$(NF-1); $NF=""; $(NF-1)=""; name=$0; sub(/\(.*/,"",name); printf "%-44s %-8s %6d\n", name, op, n }' | sort
