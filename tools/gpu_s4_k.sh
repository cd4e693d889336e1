#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_topk_tc2 -s 1 -c 1 -o gpurun_out/s4_eval2 python tools/eval_bench.py 37888 1000000 50 > gpurun_out/s4_eval2.log 2>&1
tail -3 gpurun_out/s4_eval2.log
NNCF_EVAL_DBG=2 timeout 900 ncu --set full --clock-control none -k regex:eval_topk_tc2 -s 1 -c 1 -o gpurun_out/s4_eval2_dbg2 python tools/eval_bench.py 37888 1000000 50 > gpurun_out/s4_eval2_dbg2.log 2>&1
tail -3 gpurun_out/s4_eval2_dbg2.log
ls -la gpurun_out/*.ncu-rep
