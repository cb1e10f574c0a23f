#!/bin/bash
# usage: build_variant.sh NAME search|final|final,capi "-DX264CU_MINB_SEARCH=6 ..."  -> x264-i386pic_b200/build/var/lib_NAME.so
# (tuning builds only: the named kernel files recompiled with extra defines, the rest taken from the in-tree build)
set -e
cd "$(dirname "$0")/../../x264-i386pic_b200"
mkdir -p build/var
FILES=$(echo $2 | tr ',' ' ')
for f in $FILES; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $3 -c csrc/$f.cu -o build/var/${f}_$1.o
done
OBJS=""
for o in lookahead search final aq_weights mbtree primitives primitives_mc capi stream; do
  if echo " $FILES " | grep -q " $o "; then OBJS="$OBJS build/var/${o}_$1.o"; else OBJS="$OBJS build/$o.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/var/lib_$1.so $OBJS -cudart shared -Xlinker -rpath,/usr/local/cuda/lib64
