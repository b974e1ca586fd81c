#!/bin/sh
# developer helper (build container): compile ONE source with extra -D flags and link it with the cached objects of the regular build
# into tiny-nerf-pytorch_b200/_variants/libtnerf_<name>.so (git-ignored, travels with gpurun) for A/B runs with TNERF_LIB=...
#   tools/build_variant.sh <name> <source.cu> [-DFLAG ...]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME="$1"; SRC="$2"; shift 2
CS="$ROOT/tiny-nerf-pytorch_b200/csrc"
OUT="$ROOT/tiny-nerf-pytorch_b200/_variants"
mkdir -p "$OUT" "$CS/_build/var_$NAME"
BASE="$(basename "$SRC" .cu)"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c -o "$CS/_build/var_$NAME/$BASE.o" "$CS/$SRC"
OBJS=""
for o in "$CS"/_build/*.o; do
  if [ "$(basename "$o")" = "$BASE.o" ]; then OBJS="$OBJS $CS/_build/var_$NAME/$BASE.o"; else OBJS="$OBJS $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libtnerf_$NAME.so" $OBJS
echo "built $OUT/libtnerf_$NAME.so"
