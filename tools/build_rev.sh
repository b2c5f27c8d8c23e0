#!/bin/bash
# tools/build_rev.sh REV NAME: build the library of git revision REV into build/NAME.so (for A/B runs on one box:
# RZ_B200_LIB=build/NAME.so python tools/run_configs.py)
set -e
cd "$(dirname "$0")/.."
rev=$1; name=$2
tmp=$(mktemp -d)
mkdir -p $tmp/rusterizer_b200/csrc $tmp/include build
for f in $(git ls-tree --name-only $rev rusterizer_b200/csrc/); do git show $rev:$f > $tmp/$f; done
git show $rev:include/rz.h > $tmp/include/rz.h
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true \
  -ftz=false -Xcompiler -fPIC,-ffp-contract=off -shared -o build/$name.so $tmp/rusterizer_b200/csrc/rz_api.cu
rm -rf $tmp
echo built build/$name.so from $rev
