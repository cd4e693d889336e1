#!/bin/bash
for v in 22 31 32; do echo "== stages/gbufs $v"; for a in "512 1" "512 37" "512 74" "2048 4"; do timeout 60 ./build/score_bench_$v $a | head -3 | grep -v drops; done; done
