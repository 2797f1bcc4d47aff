#!/bin/bash
# builds build_exp/lib_<tag>.so with extra nvcc flags for ONE source file (kernel experiments); the other objects are reused
#   tools/build_variant.sh <tag> <source.cu> <extra nvcc flags...>;  run with NEMO_B200_LIB=build_exp/lib_<tag>.so
set -e
tag=$1; src=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/build_exp
base=${src%.cu}
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -c $root/nemo_b200/csrc/$src -o $root/build_exp/${base}_$tag.o
objs=""
for o in $root/nemo_b200/csrc/build/*.o; do
  if [ "$(basename $o)" != "$base.o" ]; then objs="$objs $o"; fi
done
nvcc -shared -o $root/build_exp/lib_$tag.so $objs $root/build_exp/${base}_$tag.o -gencode arch=compute_100a,code=sm_100a
echo built build_exp/lib_$tag.so
