#!/bin/bash
# tools/build_variant.sh NAME [-D... nvcc flags]: builds build_variants/libp2de_NAME.so (A/B builds for tools/ab_variants.sh)
set -e
name=$1; shift
mkdir -p build_variants
cd p2de_b200/csrc
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -split-compile=0 -diag-suppress 177 \
  -Xcompiler -fPIC -shared -ldl "$@" -o ../../build_variants/libp2de_$name.so capi.cu > ../../build_variants/$name.log 2>&1
echo "built $name"
