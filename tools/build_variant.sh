#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFOO=1 ..."  -> build/variants/NAME.so (fb_kernels_fast.cu recompiled with the defines)
set -e
cd "$(dirname "$0")/../fuzzyblue_b200/csrc"
make -s -j4
mkdir -p ../../build/variants
HOSTCXX=$(command -v /usr/bin/g++ || echo g++)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin $HOSTCXX -O3 -std=c++17 -lineinfo \
  -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off --expt-relaxed-constexpr $2 -Xptxas -v -c fb_kernels_fast.cu -o ../../build/variants/$1.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin $HOSTCXX -shared -o ../../build/variants/$1.so \
  fb_api.o fb_sharded.o fb_kernels_ref.o ../../build/variants/$1.o fb_render.o fb_peak.o -cudart static -ldl
echo built build/variants/$1.so
