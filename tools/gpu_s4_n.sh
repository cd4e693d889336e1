#!/bin/bash
for t in 0 1; do echo "== touch=$t"; timeout 60 ./build/score_bench 512 37 1 $t | grep -v "stamp 1[0-9]\|stamp 2[0-9]\|stamp 3\|stamp  [89]"; done
mkdir -p gpurun_out; rm -f gpurun_out/tl.txt
NNCF_FUSE_SGD=0 NNCF_TIMELINE=gpurun_out/tl.txt timeout 600 python bench.py --steps 3000 --warmup 100 --no-eval --cpu-steps 20 > /dev/null 2>&1
echo "== in situ, NNCF_FUSE_SGD=0"; python tools/timeline.py gpurun_out/tl.txt 200
