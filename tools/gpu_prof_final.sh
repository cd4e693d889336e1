#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the dominant kernels (1 GPU)
mkdir -p gpurun_out
B="python bench.py --steps 40 --warmup 10 --no-eval --cpu-steps 3 --links 4000000"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gather_rows|score_grad|finalize|loss_out' -s 150 -c 160 --csv \
   --log-file gpurun_out/r01_launches_train.csv $B > gpurun_out/r01_launches_train.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'score_grad_tc' -s 30 -c 1 -o gpurun_out/r01_score $B > gpurun_out/r01_score.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'gather_rows_vec' -s 30 -c 1 -o gpurun_out/r01_gather $B > gpurun_out/r01_gather.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'finalize_vec' -s 30 -c 1 -o gpurun_out/r01_finalize $B > gpurun_out/r01_finalize.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:eval_topk_tc -c 1 -o gpurun_out/r01_eval python tools/eval_bench.py 37888 1000000 50 > gpurun_out/r01_eval.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'eval_topk|rows_to_img|topk_merge' --csv \
   --log-file gpurun_out/r01_launches_eval.csv python tools/eval_bench.py 37888 1000000 50 > gpurun_out/r01_launches_eval.log 2>&1
ls -la gpurun_out | grep r01
