#!/bin/bash
# round 2, call I: parity (folded Adam, tightened bf16 checks), Adam timings, convergence table v2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_multi.py -x -q 2>&1 | tail -6
CB="python tools/config_bench.py neg_shared skip-gram 512 128"
echo "== adam fold R=37"; timeout 120 $CB 37 1000 adam ureg 2>&1 | tail -1
echo "== adam nofold R=37"; NNCF_ADAM_FOLD=0 timeout 120 $CB 37 1000 adam ureg 2>&1 | tail -1
echo "== adam fold R=1"; timeout 120 $CB 1 3000 adam ureg 2>&1 | tail -1
echo "== adam nofold R=1"; NNCF_ADAM_FOLD=0 timeout 120 $CB 1 3000 adam ureg 2>&1 | tail -1
echo "== sgd R=37"; timeout 120 $CB 37 3000 ureg 2>&1 | tail -1
timeout 1500 python tools/convergence.py --epochs 40 --out gpurun_out/r02_convergence.md > gpurun_out/r02i_convergence.log 2>&1; tail -75 gpurun_out/r02i_convergence.log
