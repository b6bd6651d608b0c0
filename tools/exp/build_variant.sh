#!/bin/bash
# A/B builds: tools/exp/build_variant.sh <suffix> <source.cu> <-Dflag ...>  ->  3d_multiview_reg_b200/liblmpcr_b200_<suffix>.so
# (the regular objects of build/ with ONE source recompiled with extra flags; run the regular build first; select with LMPCR_B200_LIB)
set -e
cd "$(dirname "$0")/../../3d_multiview_reg_b200"
sfx=$1; src=$2; shift 2
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c csrc/$src -o build/${src%.cu}_$sfx.o 2>/dev/null
objs=""
for o in build/*.o; do
  b=$(basename $o .o)
  case $b in *_*_*|${src%.cu}) continue;; esac
  [[ $b == *_$sfx ]] && continue
  [[ $b =~ _[a-z0-9]+$ && ! -f csrc/$b.cu ]] && continue
  objs="$objs $o"
done
nvcc -shared -o liblmpcr_b200_$sfx.so $objs build/${src%.cu}_$sfx.o -lcudart_static -lpthread -ldl -lrt
ls -la liblmpcr_b200_$sfx.so
