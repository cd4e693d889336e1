#!/bin/bash
# Developer tool: builds libnncf_b200 variants with parts of the score kernel switched off (NNCF_ABLATE bits, see
# csrc/score_tc.cuh) into build/ablate/, to be timed in the real launch chain with
#   NNCF_LIB_PATH=build/ablate/libnncf_ab<N>.so python tools/config_bench.py neg_shared skip-gram 512 128 37 2000
# usage: tools/ablate.sh 1 2 4 ...   (run ./build.sh first: the other objects are reused)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ablate
OTHERS=$(ls build/*.o | grep -v score_tc_nsub2.o)
for n in "$@"; do
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DNNCF_ABLATE=$n -c nncf_b200/csrc/score_tc_nsub2.cu -o build/ablate/score_ab$n.o \
    && nvcc -shared -o build/ablate/libnncf_ab$n.so $OTHERS build/ablate/score_ab$n.o -Xcompiler -fPIC -lcudart ) &
done
wait
ls -la build/ablate/*.so
