#!/bin/bash
# round 2, call K: graph-captured mean-pool tower step (parity + timing), convergence v3 with the medium-scale section
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_end_to_end.py -x -q 2>&1 | tail -6
for sch in neg_shared group_neg_shared; do for gmode in 1 0; do timeout 300 python tools/tower_bench.py $sch $gmode 2>&1 | tail -1; done; done
timeout 1500 python tools/convergence.py --epochs 40 --out gpurun_out/r02_convergence.md > gpurun_out/r02k_convergence.log 2>&1; tail -22 gpurun_out/r02k_convergence.log
