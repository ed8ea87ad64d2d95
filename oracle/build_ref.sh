#!/bin/sh
# build_ref.sh -- compile the REFERENCE's own gs path into oracle/_ref/libexags_ref.so.
#
# TEST INFRASTRUCTURE ONLY.  Sources are compiled where they lie under
# /root/reference/src; the five .upc files (and the two headers that carry UPC
# qualifiers) go through strip_upc.py into a temporary directory that is
# removed afterwards.  Only the built .so lands in oracle/_ref/ (git-ignored,
# but shipped to the GPU box by gpurun).  Needs: gcc, python3.  The reference's
# own build system (autotools + a UPC compiler + libcheck) is not used -- none
# of them exists in this image.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${REF_SRC:-/root/reference/src}
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "build_ref: $REF not found"; exit 1; }
TMP=$(mktemp -d /tmp/exags_ref.XXXXXX)
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT"
for f in gs comm crystal sarray_transfer fail; do
  python3 "$HERE/strip_upc.py" "$REF/$f.upc" "$TMP/$f.c"
done
for h in comm.h crystal.h; do
  python3 "$HERE/strip_upc.py" "$REF/$h" "$TMP/$h"
done
# Nek flavour: 64-bit global ids (configure.ac:57-61) so ids are long long
FLAGS="-DPREFIX=gslib_ -DFPREFIX=fgslib_ -DGLOBAL_LONG_LONG -DUSE_NAIVE_BLAS -DUNDERSCORE"
gcc -std=gnu99 -O2 -g -fPIC -shared -w $FLAGS \
    -I"$TMP" -I"$HERE/shim" -I"$REF" \
    "$TMP/gs.c" "$TMP/comm.c" "$TMP/crystal.c" "$TMP/sarray_transfer.c" "$TMP/fail.c" \
    "$REF/gs_local.c" "$REF/sort.c" "$REF/sarray_sort.c" "$REF/tensor.c" \
    "$HERE/shim/upc_shim.c" "$HERE/ref_driver.c" \
    -o "$OUT/libexags_ref.so" -lm
echo "build_ref: wrote $OUT/libexags_ref.so"
