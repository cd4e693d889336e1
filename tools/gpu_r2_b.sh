#!/bin/bash
# round 2, call B: in-kernel clock stamps of the score kernel (stand-alone launches, touched images), base + ablations
mkdir -p gpurun_out
for n in 0 1 2 12 32 63; do
  echo "=== ablate $n  (R=37 fuse=1 touch=1)"
  timeout 60 ./build/ablate/score_bench_ab$n 512 37 1 1 2>&1 | grep -v "^    cta" | head -60
done
echo "=== ablate 0 R=1"; timeout 60 ./build/ablate/score_bench_ab0 512 1 1 1 2>&1 | head -50
NNCF_DUMP_CTAS=1 timeout 60 ./build/ablate/score_bench_ab0 512 37 1 1 > gpurun_out/r02b_ctas_ab0.txt 2>&1
NNCF_DUMP_CTAS=1 timeout 60 ./build/ablate/score_bench_ab63 512 37 1 1 > gpurun_out/r02b_ctas_ab63.txt 2>&1
