#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle/).
# Builds, into oracle/_ref/:
#   stock_K<k>_M<m>_C<c>_R<r>   the reference program as shipped (makefile:5 flags), defines patched in the gcc input stream
#   dropin_K<k>_M<m>_C<c>_R<r>  the same program with process_read / prune_data (and getval/getbp/getscore) bound to
#                               libgbin.so: the reference's own definitions are renamed away with -D, main comes from
#                               oracle/dropin_main.c, everything downstream of the hot path is unmodified reference code
# usage: oracle/build_dropin.sh K M CUTOFF READ_LENGTH
set -euo pipefail
K=$1; M=$2; C=$3; R=$4
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(dirname "$HERE")
OUT="$HERE/_ref"
mkdir -p "$OUT"
tag="K${K}_M${M}_C${C}_R${R}"
if [ ! -f "$REF/binning.c" ]; then
  echo "build_dropin: $REF/binning.c not present - keeping prebuilt binaries" >&2
  [ -x "$OUT/dropin_$tag" ] && exit 0 || exit 3
fi
PV='const int power_val[] = {1, 4, 16, 64, 256, 1024, 4096, 16384, 65536, 262144, 1048576, 4194304, 16777216, 67108864, 268435456, 1073741824};'
patch() {
  sed -e "s/^#define MMER_SIZE .*/#define MMER_SIZE $M/" -e "s/^#define KMER_SIZE .*/#define KMER_SIZE $K/" \
      -e "s/^#define ABUNDANCE_CUTOFF .*/#define ABUNDANCE_CUTOFF $C/" -e "s/^#define READ_LENGTH .*/#define READ_LENGTH $R/" \
      -e "s/^const int power_val\[\] = .*/$PV/" "$REF/binning.c"
}
# the reference as shipped (gcc -g, makefile:2)
patch | gcc -g -w -I"$REF" -x c - -x c "$REF/zhash.c" "$REF/llist.c" -o "$OUT/stock_$tag"
# the drop-in: reference TU with its hot-path definitions renamed away
patch | gcc -g -w -c -I"$REF" -Dmain=ref_unused_main -Dprocess_read=ref_unused_process_read -Dprune_data=ref_unused_prune_data \
    -Dprune_kmers=ref_unused_prune_kmers -Dgetval=ref_getval -Dgetbp=ref_getbp -Dgetscore=ref_getscore -x c - -o "$OUT/binning_$tag.o"
gcc -g -w -I"$REF" -I"$ROOT/include" -DDROPIN_K=$K -DDROPIN_M=$M -DDROPIN_CUTOFF=$C -DDROPIN_READ_LENGTH=$R \
    "$HERE/dropin_main.c" "$OUT/binning_$tag.o" "$REF/zhash.c" "$REF/llist.c" \
    -L"$ROOT/genome-assembly_b200" -lgbin -Wl,-rpath,'$ORIGIN/../../genome-assembly_b200' -o "$OUT/dropin_$tag"
rm -f "$OUT/binning_$tag.o"
echo "$OUT/stock_$tag $OUT/dropin_$tag"
