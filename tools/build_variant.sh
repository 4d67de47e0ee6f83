#!/bin/bash
# Build a tuning variant of the library next to the shipped one: tools/build_variant.sh NAME [-DFLAG=...]...
# -> hippopt_b200/variants/libhb_NAME.so (git-ignored, travels to the GPU box); select it with HIPPOPT_B200_LIB.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p hippopt_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -o hippopt_b200/variants/libhb_$name.so hippopt_b200/csrc/hippopt_b200.cu
echo built hippopt_b200/variants/libhb_$name.so "$@"
