#!/usr/bin/env bash
# Build the UNMODIFIED reference (c-kzg-4844 @ b7e4098 + vendored blst @ e7f90de) into
# oracle/_ref/libckzg_ref.so, compiling the sources where they lie under /root/reference.
# No reference source is copied into this repo; oracle/_ref/ is git-ignored (outputs only).
#
#   blst:  one C unity file (blst/src/server.c) + one assembly unity file (blst/build/assembly.S);
#          -D__BLST_PORTABLE__ = runtime dispatch to the ADX/mulx path (as SURVEY.md App. C).
#   ckzg:  one C unity file (src/ckzg.c).
#
# ORACLE / TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's
# cpu_baseline / --impl reference legs as the checker / CPU baseline.  Never linked by the product.
set -euo pipefail
REF=${REF:-/root/reference}
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) -- keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"
CC=${CC:-gcc}
$CC -O2 -fPIC -D__BLST_PORTABLE__ -fno-builtin -c "$REF/blst/src/server.c" -o "$OUT/blst_server.o"
$CC -O2 -fPIC -D__BLST_PORTABLE__ -c "$REF/blst/build/assembly.S" -o "$OUT/blst_asm.o"
$CC -O2 -fPIC -shared -I"$REF/src" -I"$REF/blst/bindings" -o "$OUT/libckzg_ref.so" \
    "$REF/src/ckzg.c" "$OUT/blst_server.o" "$OUT/blst_asm.o"
rm -f "$OUT/blst_server.o" "$OUT/blst_asm.o"
echo "built $OUT/libckzg_ref.so"
