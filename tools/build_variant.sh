#!/bin/bash
# usage: tools/build_variant.sh NAME SRC.cu "FLAGS"  -> pst/libpst_b200_NAME.so (the other objects come from build/)
set -e
cd "$(dirname "$0")/../protein-structure-tokenizer_b200"
name=$1; src=$2; flags=$3
base=$(basename $src .cu)
mkdir -p build_var
extra=""
[ "$base" = "featurize" ] && extra="-fmad=false"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I ../include -I csrc $extra $flags -c csrc/$src -o build_var/${base}_$name.o 2>/dev/null
objs=""
for o in build/*.o; do [ "$(basename $o .o)" = "$base" ] && objs="$objs build_var/${base}_$name.o" || objs="$objs $o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o pst/libpst_b200_$name.so $objs
echo pst/libpst_b200_$name.so
