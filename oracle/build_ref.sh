#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle/).
# Builds the UNMODIFIED reference hot path (binning.c + zhash.c + llist.c) from where the sources lie
# under $REF (default /root/reference) into oracle/_ref/ref_K<k>_M<m>_C<cutoff>_R<read_length>[_O0].
# The four compile-time #defines (binning.c:10-13) are not #ifndef-guarded, so they are patched in the
# *stream* fed to gcc (sed | gcc -x c -); power_val (binning.c:17) is extended to 16 entries because
# MMER_SIZE > 8 reads it out of bounds otherwise.  Nothing is copied into the repo.
#
# usage: oracle/build_ref.sh K M CUTOFF READ_LENGTH [O0|O2]
set -euo pipefail
K=$1; M=$2; C=$3; R=$4; OPT=${5:-O2}
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/_ref"
mkdir -p "$OUT"
name="ref_K${K}_M${M}_C${C}_R${R}"
flags="-O2"
if [ "$OPT" = "O0" ]; then name="${name}_O0"; flags="-g"; fi   # makefile:2 builds with -g (i.e. -O0)
if [ ! -f "$REF/binning.c" ]; then
  echo "build_ref: $REF/binning.c not present (GPU box?) - keeping prebuilt $OUT/$name" >&2
  [ -x "$OUT/$name" ] && exit 0 || exit 3
fi
PV='const int power_val[] = {1, 4, 16, 64, 256, 1024, 4096, 16384, 65536, 262144, 1048576, 4194304, 16777216, 67108864, 268435456, 1073741824};'
{
  sed -e "s/^#define MMER_SIZE .*/#define MMER_SIZE $M/" \
      -e "s/^#define KMER_SIZE .*/#define KMER_SIZE $K/" \
      -e "s/^#define ABUNDANCE_CUTOFF .*/#define ABUNDANCE_CUTOFF $C/" \
      -e "s/^#define READ_LENGTH .*/#define READ_LENGTH $R/" \
      -e "s/^const int power_val\[\] = .*/$PV/" \
      "$REF/binning.c"
  cat "$HERE/ref_harness_main.c"
} | gcc $flags -w -Dmain=ref_main -I"$REF" -x c - -x c "$REF/zhash.c" "$REF/llist.c" -o "$OUT/$name"
echo "$OUT/$name"
