#!/bin/bash
# tools/build_variant.sh NAME [-DMACRO=VALUE ...]: build an experimental librz variant into build/NAME.so
# (build/ is git-ignored but travels to the GPU box); select it with RZ_B200_LIB=build/NAME.so.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true \
  -ftz=false -Xcompiler -fPIC,-ffp-contract=off -shared "$@" -o build/$name.so rusterizer_b200/csrc/rz_api.cu
echo built build/$name.so
