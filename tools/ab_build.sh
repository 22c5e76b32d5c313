#!/bin/bash
# Build the kernels of a git revision (default HEAD) into open-headstage_b200/libohs_cuda_<tag>.so for A/B runs with
# OHS_LIB_OVERRIDE (the library is git-ignored and travels to the GPU box with the snapshot).
set -e
rev=${1:-HEAD}; tag=${2:-prev}
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p $tmp/open-headstage_b200/csrc $tmp/include
for f in $(git -C $root ls-tree -r --name-only $rev open-headstage_b200/csrc include); do git -C $root show $rev:$f > $tmp/$f; done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -I$tmp/include \
  -o $root/open-headstage_b200/libohs_cuda_$tag.so $tmp/open-headstage_b200/csrc/ohs_api.cu
rm -rf $tmp; echo built libohs_cuda_$tag.so from $rev
