#!/usr/bin/env bash
# Static evidence from the built library (no GPU needed): per-kernel registers / spills / shared memory from the cubin
# (cuobjdump -res-usage) and counts of the SASS mnemonics that identify the Blackwell paths (B200_PROFILING.md):
#   UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG/UBLKCP = TMA, HMMA = legacy mma.sync.
# Usage: bash profiles/tools/static_evidence.sh > profiles/r01_static_sass.txt
set -e
LIB=neural-network-from-scratch_b200/lib/libnfb200.so
echo "# $(date -u +%F) $(nvcc --version | tail -1)"
echo "# library: $LIB ($(stat -c %s $LIB) bytes), arch: $(cuobjdump -lelf $LIB | head -3 | tr '\n' ' ')"
echo
echo "## SASS mnemonic counts per kernel (kernels with none of these are plain SIMT bandwidth kernels)"
cuobjdump -sass $LIB | awk '
  /Function :/ { fn=$3 }
  { for (i=1;i<=NF;i++) { t=$i; sub(/\..*/,"",t);
      if (t ~ /^UTC[A-Z]*MMA$/) c[fn,"UTCxMMA"]++;
      else if (t=="LDTM") c[fn,"LDTM"]++; else if (t=="STTM") c[fn,"STTM"]++;
      else if (t=="UTMALDG") c[fn,"UTMALDG"]++; else if (t=="UTMASTG") c[fn,"UTMASTG"]++; else if (t=="UTMAREDG") c[fn,"UTMAREDG"]++;
      else if (t=="UBLKCP") c[fn,"UBLKCP"]++; else if (t=="HMMA") c[fn,"HMMA"]++; else if (t=="MUFU") c[fn,"MUFU"]++;
      else continue; seen[fn]=1 } }
  END { n=split("UTCxMMA LDTM STTM UTMALDG UTMASTG UTMAREDG UBLKCP HMMA MUFU",k," ");
        for (f in seen) { line=""; tc=0; for (j=1;j<=n;j++) { v=c[f,k[j]]+0; if (j<=8) tc+=v; if (v) line=line sprintf(" %s=%d",k[j],v) }
          if (tc) print f ":" line } }' | sort | while read -r line; do
    fn=${line%%:*}; echo "$(echo "$fn" | c++filt | cut -c1-110):${line#*:}"; done
echo
echo "## registers / spills / shared memory per kernel (cuobjdump -res-usage)"
cuobjdump -res-usage $LIB | awk '/Function/ { fn=$2; sub(/:$/,"",fn) } /REG:/ { print fn " " $0 }' | while read -r fn rest; do
  echo "$(echo "$fn" | c++filt | cut -c1-100)  $rest"; done | sort
