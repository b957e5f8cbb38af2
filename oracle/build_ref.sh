#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference CPU tools (Kmer-db 2.3.1, LZ-ANI 1.2.3)
# straight from the read-only sources under /root/reference/3rd_party into oracle/_ref/ (git-ignored).
# No reference source is copied into this repository: g++ is pointed at the files where they lie.
# The reference's own build system (refresh.mk + cmake'd zlib-ng + mimalloc) is NOT run; instead
#   * <zlib-ng/zlib.h> is satisfied by a one-line shim that includes the system <zlib.h> (zlib-ng was used by
#     the reference in ZLIB_COMPAT mode, i.e. through the plain zlib API), and
#   * <mimalloc-new-delete.h> (a malloc override, no functional effect) by an empty shim.
# Used by: tests/ (parity goldens), bench.py cpu_baseline / --impl reference.  Never by the product path.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${VCLUST_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$REF/3rd_party/lz-ani/src" ]; then
  echo "build_ref: $REF not present; keeping prebuilt binaries in $OUT (if any)" >&2
  exit 0
fi
if [ -x "$OUT/kmer-db" ] && [ -x "$OUT/lz-ani" ] && [ "${FORCE:-0}" != 1 ]; then
  exit 0
fi
mkdir -p "$OUT/shim/zlib-ng" "$OUT/obj"
echo '#include <zlib.h>' > "$OUT/shim/zlib-ng/zlib.h"
: > "$OUT/shim/mimalloc-new-delete.h"
# x86-64-v3 (AVX2) rather than -march=native: the binaries travel to the GPU box, whose host CPU may differ.
CXXFLAGS="-O3 -std=c++20 -march=x86-64-v3 -DARCH_X64 -w -I $OUT/shim"

L="$REF/3rd_party/lz-ani"
K="$REF/3rd_party/kmer-db"
compile() { # $1 = tag, $2 = include root, rest = sources
  local tag="$1" inc="$2"; shift 2
  printf '%s\n' "$@" | xargs -P "$JOBS" -I{} sh -c \
    "g++ $CXXFLAGS -I $inc -c {} -o $OUT/obj/${tag}_\$(basename {}).o"
}
compile lz "$L/libs" "$L"/src/*.cpp
g++ -o "$OUT/lz-ani" "$OUT"/obj/lz_*.o -lz -lpthread
compile km "$K/libs" "$K"/src/*.cpp "$K"/src/kmc_api/*.cpp "$K/src/simd/row_add_avx.cpp" "$K/src/simd/row_add_avx2.cpp"
g++ -o "$OUT/kmer-db" "$OUT"/obj/km_*.o -lz -lpthread
rm -rf "$OUT/obj"
echo "build_ref: built $OUT/kmer-db and $OUT/lz-ani" >&2
