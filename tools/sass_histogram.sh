#!/bin/bash
# SASS evidence per kernel of libnpvc_b200.so (no GPU needed): instruction count and the opcodes that prove what the
# kernel is -- UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA tensor load / store), UBLKCP (bulk copy), LDTM (tcgen05.ld),
# UTCBAR (tcgen05.commit), SYNCS (mbarrier), STG.*.256 / LDG.*.256 (32-byte accesses), RED (vector reductions),
# FFMA2 (packed fp32), HMMA / IMMA (would be mma.sync: expected 0).
#   bash tools/sass_histogram.sh > profiles/r2_sass_histogram.txt
SO=${1:-vae_npvc_b200/libnpvc_b200.so}
echo "# $(basename $SO)  md5 $(md5sum $SO | cut -c1-12)  $(date -u +%F)"
printf "%-72s %7s %8s %8s %7s %6s %6s %6s %7s %7s %5s %6s %5s\n" kernel instrs UTCHMMA UTMALDG UBLKCP LDTM UTCBAR SYNCS STG256 LDG256 RED FFMA2 HMMA
cuobjdump -sass "$SO" 2>/dev/null | awk '
  /Function :/ { if (name != "") out(); name = $3; n = 0; delete c; next }
  /^[ \t]+\/\*[0-9a-f]{4,6}\*\// { n++; op = $2; if (op ~ /^@/) op = $3;
     if (op ~ /^UTCHMMA/) c["m"]++; if (op ~ /^UTMALDG/) c["t"]++; if (op ~ /^UBLKCP/) c["b"]++; if (op ~ /^LDTM/) c["l"]++;
     if (op ~ /^UTCBAR/) c["u"]++; if (op ~ /^SYNCS/) c["s"]++; if (op ~ /^STG.*256/) c["g"]++; if (op ~ /^LDG.*256/) c["d"]++;
     if (op ~ /^RED/) c["r"]++; if (op ~ /^FFMA2/) c["f"]++; if (op ~ /^(HMMA|IMMA)/) c["h"]++ }
  function out() { printf "%-72s %7d %8d %8d %7d %6d %6d %6d %7d %7d %5d %6d %5d\n", substr(name, 1, 72), n, c["m"], c["t"], c["b"], c["l"], c["u"], c["s"], c["g"], c["d"], c["r"], c["f"], c["h"] }
  END { if (name != "") out() }' | sed 's/_ZN4npvc//'
