#!/bin/bash
# round 2, call C: split sweep + folded regulariser: parity tests, then timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py -x -q 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_train_step.py 2>&1 | tail -3
CB="python tools/config_bench.py neg_shared skip-gram 512 128"
for sp in 1 2 4 8; do echo "== R=1 split $sp"; NNCF_SPLIT=$sp timeout 120 $CB 1 3000 2>&1 | tail -1; done
echo "== R=1 auto ureg"; timeout 120 $CB 1 3000 ureg 2>&1 | tail -1
echo "== R=1 auto adam ureg"; timeout 120 $CB 1 3000 adam ureg 2>&1 | tail -1
for sp in 1 4 8; do echo "== R=1 adam split $sp"; NNCF_SPLIT=$sp timeout 120 $CB 1 3000 adam 2>&1 | tail -1; done
echo "== R=4 auto"; timeout 120 $CB 4 3000 2>&1 | tail -1
echo "== R=4 split1"; NNCF_SPLIT=1 timeout 120 $CB 4 3000 2>&1 | tail -1
echo "== R=9 auto"; timeout 120 $CB 9 3000 2>&1 | tail -1
echo "== R=18 auto"; timeout 120 $CB 18 3000 2>&1 | tail -1
echo "== R=37"; timeout 120 $CB 37 3000 2>&1 | tail -1
echo "== R=37 ureg"; timeout 120 $CB 37 3000 ureg 2>&1 | tail -1
echo "== R=37 adam ureg"; timeout 120 $CB 37 1000 adam ureg 2>&1 | tail -1
