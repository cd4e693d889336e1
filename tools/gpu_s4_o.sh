#!/bin/bash
for e in 1.0 0.5 0.0; do echo "== exponent scale $e (0 = uniform ids, no hot rows)"; timeout 60 ./build/score_bench 512 37 1 0 $e | grep "stamp  [3-6]\|CTAs\|per launch"; done
