#!/bin/sh
# Kernel experiments: build the product library with other flags for ONE translation unit into gr-ieee80211_b200/lib/variants/
#   tools/build_variant.sh NAME FILE.cu "EXTRA NVCC FLAGS"      ->  lib/variants/libc80211b200_NAME.so
# and select it at run time with C8B_LIB=<path> (see _cabi.py).  The default build is untouched.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/gr-ieee80211_b200/csrc
OUT=$ROOT/gr-ieee80211_b200/lib/variants
NAME=$1; FILE=$2; EXTRA=$3
mkdir -p "$OUT" "$SRC/build/var_$NAME"
ARCH="-gencode arch=compute_100a,code=sm_100a"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC --fmad=true $EXTRA -c "$SRC/$FILE" -o "$SRC/build/var_$NAME/${FILE%.cu}.o"
OBJS=""
for o in "$SRC"/build/*.o; do
  b=$(basename "$o")
  if [ "$b" = "${FILE%.cu}.o" ]; then OBJS="$OBJS $SRC/build/var_$NAME/$b"; else OBJS="$OBJS $o"; fi
done
/usr/local/cuda/bin/nvcc $ARCH -shared -o "$OUT/libc80211b200_$NAME.so" $OBJS
echo "$OUT/libc80211b200_$NAME.so"
