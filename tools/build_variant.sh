#!/bin/bash
# Build libmfb200_<tag>.so with extra -D flags on fast.cu (kernel experiments; load with MFB_LIB=...).
# usage: tools/build_variant.sh <tag> [-DNAME=VALUE ...]      (optional: SRC=<other fast.cu>)
set -e
cd "$(dirname "$0")/../microstructure_fingerprinting_b200"
tag=$1; shift
src=${SRC:-csrc/fast.cu}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Icsrc "$@" -c "$src" -o /tmp/fast_$tag.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libmfb200_$tag.so csrc/exact.o /tmp/fast_$tag.o csrc/mc.o csrc/api.o -lcudart
echo built libmfb200_$tag.so
