#!/usr/bin/env bash
# Counts, per embedded cubin of libkektordb_gpu.so, the SASS mnemonics that prove the Blackwell paths are
# the ones compiled in (B200_PROFILING.md): UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor loads), LDTM (tcgen05.ld),
# UTCBAR (tcgen05.commit -> mbarrier), UBLKCP (cp.async.bulk 1-D bulk copies), SYNCS (mbarrier), IDP.4A (dp4a), FFMA2 (packed f32 FMA), ELECT (elect.sync).
#   bash profiles/tools/sass_summary.sh > profiles/sass_summary.txt
set -euo pipefail
cd "$(dirname "$0")/../.."
LIB=kektordb_b200/libkektordb_gpu.so
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
( cd "$TMP" && cuobjdump -xelf all "$OLDPWD/$LIB" >/dev/null )
echo "# SASS evidence per cubin of $LIB ($(nvcc --version | grep release | sed 's/.*release //'))"
echo "# built from HEAD $(git rev-parse --short HEAD) (+ working tree); arch $(cuobjdump -lelf $LIB | head -1 | sed 's/.*\.\(sm_[0-9a-z]*\)\..*/\1/')"
printf "%-28s %9s %8s %6s %7s %7s %7s %7s %7s %7s %7s\n" cubin instrs UTCHMMA LDTM UTMALDG UTCBAR UBLKCP SYNCS IDP.4A FFMA2 ELECT
for f in "$TMP"/*.cubin; do
  s=$(cuobjdump -sass "$f")
  n() { grep -c "$1" <<<"$s" || true; }
  name=$(basename "$f" | sed 's/\.sm_100a\.cubin//; s/^[^.]*\.[0-9]*\.//')
  printf "%-28s %9s %8s %6s %7s %7s %7s %7s %7s %7s %7s\n" "$name" "$(grep -c ';' <<<"$s")" "$(n UTCHMMA)" "$(n LDTM)" "$(n UTMALDG)" "$(n UTCBAR)" "$(n UBLKCP)" "$(n 'SYNCS')" "$(n 'IDP.4A')" "$(n 'FFMA2')" "$(n 'ELECT')"
done
echo
echo "# kernels per cubin (cuobjdump -sass | grep Function)"
for f in "$TMP"/*.cubin; do
  name=$(basename "$f" | sed 's/\.sm_100a\.cubin//; s/^[^.]*\.[0-9]*\.//')
  echo "$name: $(cuobjdump -sass "$f" | grep -c 'Function :') kernels"
done
