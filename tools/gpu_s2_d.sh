#!/bin/bash
mkdir -p gpurun_out
for v in 32 22; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:score_grad -s 5 -c 1 -f -o gpurun_out/s2_score_$v ./build/score_bench_$v 512 37 > gpurun_out/s2_score_$v.log 2>&1
tail -2 gpurun_out/s2_score_$v.log
done
ls -la gpurun_out/*.ncu-rep
