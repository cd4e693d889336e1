#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_end_to_end.py tests/test_gpu_eval_sampler_batch.py -x -q 2>&1 | tail -15
for sch in neg_shared group_neg_shared; do timeout 300 python tools/tower_bench.py $sch 1 2>&1 | tail -2; done
timeout 1500 python tools/convergence.py --epochs 40 --out gpurun_out/r02_convergence.md > gpurun_out/r02m_convergence.log 2>&1; tail -3 gpurun_out/r02m_convergence.log
