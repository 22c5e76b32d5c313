#!/bin/bash
# Build the kernels of a git revision (or WORK = the working tree) into open-headstage_b200/libohs_cuda_<tag>.so for A/B
# runs with OHS_LIB_OVERRIDE (the library is git-ignored and travels to the GPU box with the snapshot).
# usage: ab_build.sh <rev|WORK> <tag> [extra nvcc flags, e.g. -DOHS_EQ_WEIGHT=3]
set -e
rev=${1:-HEAD}; tag=${2:-prev}; shift 2 || true
root=$(cd "$(dirname "$0")/.." && pwd)
if [ "$rev" = WORK ]; then src=$root; else
  src=$(mktemp -d); mkdir -p $src/open-headstage_b200/csrc $src/include
  for f in $(git -C $root ls-tree -r --name-only $rev open-headstage_b200/csrc include); do git -C $root show $rev:$f > $src/$f; done
fi
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o $root/open-headstage_b200/libohs_cuda_$tag.so $src/open-headstage_b200/csrc/ohs_api.cu
[ "$rev" = WORK ] || rm -rf $src; echo built libohs_cuda_$tag.so from $rev "$@"
