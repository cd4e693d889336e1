#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 40 --warmup 10 --no-eval --cpu-steps 3 --links 4000000"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'score_grad_tc' -s 30 -c 1 -o gpurun_out/s4_score $B > gpurun_out/s4_score.log 2>&1; tail -2 gpurun_out/s4_score.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gather_rows_vec' -s 30 -c 1 -o gpurun_out/s4_gather $B > gpurun_out/s4_gather.log 2>&1; tail -2 gpurun_out/s4_gather.log
