#!/bin/bash
# Builds libnncf_b200.so (sm_100a) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
SRC="nncf_b200/csrc/api.cu nncf_b200/csrc/train_step.cu nncf_b200/csrc/score_tc_nsub1.cu nncf_b200/csrc/score_tc_nsub2.cu nncf_b200/csrc/score_tc_nsub4.cu nncf_b200/csrc/eval_topk.cu nncf_b200/csrc/sampler.cu nncf_b200/csrc/group_sampler.cu nncf_b200/csrc/batch_builder.cu nncf_b200/csrc/meanpool.cu nncf_b200/csrc/peer.cu"
mkdir -p build
OBJS=""
for f in $SRC; do
  o=build/$(basename ${f%.cu}).o
  if [ ! -f $o ] || [ $f -nt $o ] || [ nncf_b200/csrc/common.cuh -nt $o ] || [ nncf_b200/csrc/sm100.cuh -nt $o ] || [ nncf_b200/csrc/score_tc.cuh -nt $o ] || [ nncf_b200/csrc/row_kernels.cuh -nt $o ] || [ nncf_b200/csrc/sns_kernels.cuh -nt $o ] || [ nncf_b200/csrc/pairs_kernels.cuh -nt $o ] || [ nncf_b200/csrc/alias.cuh -nt $o ] || [ include/nncf_b200.h -nt $o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c $f -o $o &
  fi
  OBJS="$OBJS $o"
done
wait
nvcc -shared -o nncf_b200/libnncf_b200.so $OBJS -Xcompiler -fPIC -lcudart
echo "built nncf_b200/libnncf_b200.so"
