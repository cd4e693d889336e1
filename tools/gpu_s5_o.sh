#!/bin/bash
mkdir -p gpurun_out
NNCF_EVAL_GEN=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_topk_tc3 -s 1 -c 1 -o gpurun_out/s5o_eval3 python tools/eval_bench.py 37888 1000000 50 > gpurun_out/s5o_eval3.log 2>&1
tail -2 gpurun_out/s5o_eval3.log
