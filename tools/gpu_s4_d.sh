#!/bin/bash
for m in 0 1 2; do echo "== drain mode $m fused"; timeout 60 ./build/score_bench_m$m 512 37 1 | grep -v "stamp 1[0-9]\|stamp 2[0-9]\|stamp 3\|stamp  [89]"; done
echo "== non-fused"; timeout 60 ./build/score_bench_m0 512 37 0 | grep -v "stamp 1[0-9]\|stamp 2[0-9]\|stamp 3\|stamp  [89]"
