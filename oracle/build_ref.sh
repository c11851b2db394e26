#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the UNMODIFIED reference CPU path (GJK/cpu/openGJK.c + GJK/cpu/EPA.c) from where
# it lies under $REF (default /root/reference) into oracle/_ref/libogjk_ref_f32.so and
# libogjk_ref_f64.so.  Nothing from the reference is copied into the repository:
#   * fp32 is compiled in place (USE_32BITS is hard-defined at GJK/common.h:44);
#   * fp64 needs that one #define removed, which cannot be done from the command line, so
#     the three directories are copied to a mktemp scratch dir, line 44 is commented out with
#     sed, the scratch copy is compiled and then deleted.
# Flags: -O3 as the reference Release build (CMakeLists.txt:31-37) minus -Werror,
#   -ffp-contract=off   keep IEEE op-by-op evaluation (x86-64 baseline has no FMA anyway),
#   -include string.h   GJK/cpu/EPA.c:145 uses memset without including it.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -f "$REF/GJK/cpu/openGJK.c" ]; then
  echo "build_ref: $REF not present; keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
CFLAGS="-std=gnu11 -O3 -ffp-contract=off -fno-fast-math -fPIC -shared -fopenmp -include string.h -w -Wl,-Bsymbolic"
gcc $CFLAGS -I"$REF" "$HERE/ref_driver.c" "$REF/GJK/cpu/openGJK.c" "$REF/GJK/cpu/EPA.c" \
    -o "$OUT/libogjk_ref_f32.so" -lm
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/GJK"
cp -r "$REF/GJK/common.h" "$REF/GJK/cpu" "$TMP/GJK/"
sed -i 's,^#define USE_32BITS,// #define USE_32BITS,' "$TMP/GJK/common.h"
grep -q '^// #define USE_32BITS' "$TMP/GJK/common.h"
gcc $CFLAGS -I"$TMP" "$HERE/ref_driver.c" "$TMP/GJK/cpu/openGJK.c" "$TMP/GJK/cpu/EPA.c" \
    -o "$OUT/libogjk_ref_f64.so" -lm
echo "build_ref: wrote $OUT/libogjk_ref_f32.so $OUT/libogjk_ref_f64.so"
