#!/bin/bash
# round 2, call D: parity (train step incl. split / folded regulariser, pipelined schedule on one device), timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_multi.py tests/test_gpu_eval_sampler_batch.py -x -q 2>&1 | tail -8
CB="python tools/config_bench.py neg_shared skip-gram 512 128"
for dv in 0 1; do for sp in 2 4; do echo "== R=1 split $sp drain_vec $dv"; NNCF_DRAIN_VEC=$dv NNCF_SPLIT=$sp timeout 120 $CB 1 3000 2>&1 | tail -1; done; done
echo "== R=1 auto ureg"; timeout 120 $CB 1 3000 ureg 2>&1 | tail -1
echo "== R=1 auto adam ureg"; timeout 120 $CB 1 3000 adam ureg 2>&1 | tail -1
echo "== R=4 auto"; timeout 120 $CB 4 3000 2>&1 | tail -1
echo "== R=4 drain_vec 0"; NNCF_DRAIN_VEC=0 timeout 120 $CB 4 3000 2>&1 | tail -1
echo "== R=9 auto"; timeout 120 $CB 9 3000 2>&1 | tail -1
echo "== R=9 drain_vec 0"; NNCF_DRAIN_VEC=0 timeout 120 $CB 9 3000 2>&1 | tail -1
echo "== R=18 drain_vec 1"; NNCF_DRAIN_VEC=1 timeout 120 $CB 18 3000 2>&1 | tail -1
echo "== R=18 drain_vec 0"; NNCF_DRAIN_VEC=0 timeout 120 $CB 18 3000 2>&1 | tail -1
echo "== R=37"; timeout 120 $CB 37 3000 2>&1 | tail -1
echo "== R=37 ureg"; timeout 120 $CB 37 3000 ureg 2>&1 | tail -1
echo "== R=37 drain_vec 1"; NNCF_DRAIN_VEC=1 timeout 120 $CB 37 3000 2>&1 | tail -1
